"""CPU: the C-ABI shared library loads and exports exactly what include/stv.h declares; argument validation that needs no
GPU behaves like the reference (ValueError for bad shapes / arguments)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT/'include'/'stv.h').read_text()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(stv_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    from slowtv_monodepth_b200 import _lib as L
    lib = L.lib()
    declared = _declared()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/stv.h but not exported by libstv.so'
    assert sorted(L.exported_symbols()) == declared, 'ctypes signatures out of sync with include/stv.h'
    assert lib.stv_version() >= 100


def test_library_exports_nothing_undeclared():
    """The reverse direction: every stv_* function the shared library defines is part of the documented ABI (include/stv.h)."""
    import shutil
    import subprocess
    from slowtv_monodepth_b200 import _build
    nm = shutil.which('nm')
    if nm is None: pytest.skip('binutils nm not available')
    out = subprocess.run([nm, '-D', '--defined-only', str(_build.LIB)], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[2] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] == 'T' and ln.split()[2].startswith('stv_')})
    assert exported == _declared()


def test_invalid_arguments_are_rejected_without_a_gpu():
    from slowtv_monodepth_b200 import _lib as L
    lib = L.lib()
    cfg = L.PhotoCfg(b=1, n=2, S=9, H=16, W=16, w_ssim=0.85, w_l1=0.15, use_min=1, use_automask=1, noise_seed=0, depth_stride_s=0)
    assert lib.stv_photo_workspace_bytes(C.byref(cfg)) == 0  # S > STV_MAX_SCALES
    assert b'STV_MAX_SCALES' in lib.stv_last_error()
    cfg.S = 4
    assert lib.stv_photo_workspace_bytes(C.byref(cfg)) > 0
    rc = lib.stv_photo_fwd(C.byref(cfg), None, None, None, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == 1 and b'NULL' in lib.stv_last_error()
    with pytest.raises(ValueError): L.check(rc, 'stv_photo_fwd')
    cfg.H = 2
    assert lib.stv_photo_fwd(C.byref(cfg), None, None, None, None, None, None, None, None, None, None, None, None, 0, None) == 1
    assert lib.stv_adamw_step(None, None, None, None, 8, 0, 1e-3, .9, .999, 1e-8, 0., 1., 1, None) == 1
    assert lib.stv_disp_to_depth_fwd(1, 0, 4, 8, 8, 0.1, 100., None, None, None, None) == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from slowtv_monodepth_b200 import _lib as L
    monkeypatch.setattr(L, '_lib', None)
    monkeypatch.setattr(L, 'LIB_PATH', tmp_path/'libstv.so')
    with pytest.raises(L.StvError): L.lib()


def test_cpu_tensors_are_refused():
    import torch
    from slowtv_monodepth_b200 import _lib as L, functional as F_
    t = torch.zeros(1, 3, 8, 8)
    with pytest.raises(L.StvError):
        F_.photo_loss([torch.zeros(1, 1, 8, 8)], t, t[None], torch.eye(4)[None, None], torch.eye(4)[None])
    with pytest.raises(L.StvError):
        F_.smooth_loss([torch.zeros(1, 1, 8, 8)], t)


def test_struct_layouts_match_the_header(tmp_path):
    """Every struct of include/stv.h has the size and field offsets of its ctypes mirror (compiled with the host C compiler)."""
    import shutil
    import subprocess
    from slowtv_monodepth_b200 import _lib as L
    cc = shutil.which('gcc') or shutil.which('cc')
    if cc is None: pytest.skip('no host C compiler')
    pairs = {'stv_photo_cfg': L.PhotoCfg, 'stv_photo_src': L.PhotoSrc, 'stv_recon_cfg': L.ReconCfg, 'stv_smooth_cfg': L.SmoothCfg,
             'stv_gemm_epi': L.GemmEpi, 'stv_conv_geom': L.ConvGeom}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT/"include"/"stv.h"}"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, *_ in cls._fields_: lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path/'layout.c'
    src.write_text('\n'.join(lines))
    subprocess.run([cc, '-std=c11', '-o', str(tmp_path/'layout'), str(src)], check=True)
    out = subprocess.run([str(tmp_path/'layout')], capture_output=True, text=True, check=True).stdout
    for row in out.strip().splitlines():
        cname, size, *offs = row.split()
        cls = pairs[cname]
        assert int(size) == C.sizeof(cls), f'{cname}: sizeof {size} (header) vs {C.sizeof(cls)} (ctypes)'
        assert [int(o) for o in offs] == [getattr(cls, f[0]).offset for f in cls._fields_], f'{cname}: field offsets differ'


def test_new_entry_points_validate_their_arguments_without_a_gpu():
    from slowtv_monodepth_b200 import _lib as L
    lib = L.lib()
    cfg = L.ReconCfg(b=1, n=2, C=3, H=8, W=8, loss=7, use_min=1, use_automask=0, mask_mode=0, noise_seed=0)
    assert lib.stv_recon_ex_workspace_bytes(C.byref(cfg)) == 0 and b'bad loss' in lib.stv_last_error()
    cfg.loss, cfg.mask_mode = 0, 5
    assert lib.stv_recon_ex_workspace_bytes(C.byref(cfg)) == 0 and b'Invalid mask type' in lib.stv_last_error()
    cfg.mask_mode = 1
    assert lib.stv_recon_ex_workspace_bytes(C.byref(cfg)) > 0
    assert lib.stv_regr_fwd(16, 9, 0, None, None, None, None, None, None, 0, None) == 1          # NULL operands / bad loss
    assert lib.stv_smooth_ex_workspace_bytes(1, 3, 2, 8) == 0 and lib.stv_smooth_ex_workspace_bytes(1, 3, 8, 8) > 0
    assert lib.stv_feat_reg_workspace_bytes(1, 4, 3, 1, 8) == 0 and lib.stv_feat_reg_workspace_bytes(1, 4, 3, 8, 8) > 0
    assert lib.stv_pwreg_fwd(0, 0, 1.0, None, None, None, 0, None) == 1
