"""Build recipe for `oracle/_ref/`: the UNMODIFIED reference, compiled where it lies, as a travelling artefact.

TEST INFRASTRUCTURE. The reference (jspenmar/slowtv_monodepth) is pure Python, its tree (`/root/reference`) exists only in the
build container, and its sources must not be copied into this repository. What CAN travel to the GPU box is a *built* artefact,
exactly like a compiled `.so`: this script byte-compiles the reference's hot-path packages from the sources where they lie
(`py_compile`, no source text is written anywhere) into ONE archive `oracle/_ref/ref_build.zip` holding `src/**.pyc` in the
sourceless import layout (imported through zipimport; a single binary file, because directory trees of `*.pyc` are commonly
filtered out when a work tree is shipped), and dumps the experiment configurations the benchmarks name as parsed JSON
(`oracle/_ref/cfg.json`). `oracle/_ref/` is git-ignored (never
enters history) and not gpurun-ignored (ships with the snapshot). `__graft_entry__.build()` runs this when `/root/reference`
is present; on the GPU box the prebuilt files are used as they are.

With it, on the GPU box:
  * `tests/test_plugin_gpu.py` runs the reference's own `MonoDepthModule.step` on CUDA with and without `plugin.install()`;
  * `bench.py --impl reference` times the reference's own classes (`cpu_baseline.kind = "reference"`).
The vendored third-party trees under `src/external_libs` (DGP, MiDaS, NeWCRFs: data loaders / baselines of other papers) are not
on the path and are skipped; timm / Lightning / kornia stay the placeholders of oracle/ref_shim.py.
"""
from __future__ import annotations

import json
import py_compile
import shutil
import sys
import tempfile
import zipfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path('/root/reference')
OUT = ROOT/'oracle'/'_ref'
ARCHIVE = 'ref_build.zip'
SKIP = ('external_libs/dgp', 'external_libs/midas', 'external_libs/newcrfs')
CFGS = ['default.yaml', 'kbr/default.yaml', 'abl_learn_K/default.yaml', 'benchmark/default.yaml', 'benchmark/monodepth2_M.yaml']


def stamp() -> str:
    import hashlib
    h = hashlib.sha1()
    for f in sorted((REF/'src').rglob('*.py')):
        if any(s in f.as_posix() for s in SKIP): continue
        h.update(f.relative_to(REF).as_posix().encode()); h.update(f.read_bytes())
    return f'{sys.version_info.major}.{sys.version_info.minor}.{sys.version_info.micro}:{h.hexdigest()}'


def build(force: bool = False) -> Path | None:
    if not (REF/'src'/'core'/'trainer.py').is_file(): return OUT if (OUT/'STAMP').is_file() else None  # GPU box: prebuilt only
    st = stamp()
    if not force and (OUT/'STAMP').is_file() and (OUT/'STAMP').read_text() == st: return OUT
    if OUT.exists(): shutil.rmtree(OUT)
    OUT.mkdir(parents=True)
    n = 0
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(OUT/ARCHIVE, 'w', zipfile.ZIP_DEFLATED) as zf:
        for f in sorted((REF/'src').rglob('*.py')):
            rel = f.relative_to(REF)
            if any(s in rel.as_posix() for s in SKIP): continue
            dst = Path(tmp)/'m.pyc'
            # dfile: tracebacks keep pointing at the reference's own file:line
            py_compile.compile(str(f), cfile=str(dst), dfile=str(f), doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            zf.write(dst, rel.with_suffix('.pyc').as_posix())
            n += 1
    import yaml
    cfgs = {c: yaml.safe_load((REF/'cfg'/c).read_text()) for c in CFGS if (REF/'cfg'/c).is_file()}
    (OUT/'cfg.json').write_text(json.dumps(cfgs))
    (OUT/'STAMP').write_text(st)
    print(f'oracle/_ref: {n} modules byte-compiled from {REF}/src, {len(cfgs)} configurations parsed')
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv)
