"""Generate the golden fixtures under tests/golden/ by executing the REAL reference (from /root/reference) on CPU.

TEST INFRASTRUCTURE. Run in the build container only:  python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so these known-answer files are produced by its
own classes — handlers.image_recon / handlers.disp_smooth / ViewSynth / ReconstructionLoss / SmoothReg / T_from_AAt /
to_scaled / ops.interpolate_like — in float64 on inputs that are exactly representable after a small quantisation
(images: uint8/255, disparities: uint16/65535), so the fixture stores integers and the float64 answers only.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
warnings.filterwarnings('ignore')

from oracle import ref_shim  # noqa: E402
from slowtv_monodepth_b200 import synthetic as syn  # noqa: E402

GOLDEN = ROOT/'tests'/'golden'

CASES = {
    # name: dict(b, n, S, shape, seed, use_min, use_automask, loss_name, learn_K, pose_scale, extras)
    'c1_s1': dict(b=1, n=2, S=1, shape=(128, 192), seed=0),
    'c1_s4': dict(b=1, n=2, S=4, shape=(128, 192), seed=1),
    'ragged_n4': dict(b=2, n=4, S=2, shape=(40, 72), seed=2, learn_K=True, pose_scale=6.0),
    'mean_noauto_l1': dict(b=2, n=2, S=2, shape=(32, 48), seed=3, use_min=False, use_automask=False, loss_name='l1'),
    'behind_ties_const': dict(b=2, n=2, S=1, shape=(48, 80), seed=4, pose_scale=40.0, tie=True, const_patch=True),
    'noscale_mean_auto': dict(b=1, n=2, S=2, shape=(32, 64), seed=5, use_min=False, use_automask=True, min_depth=None, max_depth=None),
}


def quantised_inputs(cfg: dict) -> dict:
    d = syn.make_loss_inputs(cfg['b'], cfg['n'], cfg['S'], cfg['shape'], seed=cfg['seed'], learn_K=cfg.get('learn_K', False))
    q8 = lambda x: (x*255).round().clamp(0, 255).to(torch.uint8)
    q16 = lambda x: (x*65535).round().clamp(1, 65534).to(torch.int32)
    imgs, supp = q8(d['imgs']), q8(d['supp_imgs'])
    if cfg.get('tie'): supp[1] = supp[0]  # Exact ties between support frames (first index must win).
    if cfg.get('const_patch'):
        imgs[..., 8:24, 8:40] = 128; supp[..., 8:24, 8:40] = 128  # Constant region: SSIM denominators ~ C1*C2, L1 = 0.
    ps = cfg.get('pose_scale', 1.0)
    aa, t = d['aa']*ps, d['t']*ps
    if cfg.get('tie'): aa[1], t[1] = aa[0], t[0]
    if ps > 10: t[..., 2] -= 0.15  # Pushes near points behind / close to the camera (z < 0.1 clamp).
    rs = np.random.RandomState(cfg['seed'] + 100)
    H, W = cfg['shape']
    noise = rs.standard_normal((cfg['S']*cfg['b'], 1, H, W)).astype(np.float32)
    return dict(imgs_u8=imgs.numpy(), supp_u8=supp.numpy(), disps_u16=[q16(x).numpy().astype(np.uint16) for x in d['disps']],
                aa=aa.numpy().astype(np.float32), t=t.numpy().astype(np.float32), K=d['K'].numpy().astype(np.float32), noise=noise)


def dequantise(q: dict, dtype=torch.float64) -> dict:
    return dict(imgs=torch.from_numpy(q['imgs_u8']).to(dtype)/255, supp_imgs=torch.from_numpy(q['supp_u8']).to(dtype)/255,
                disps=[torch.from_numpy(x.astype(np.int32)).to(dtype)/65535 for x in q['disps_u16']],
                aa=torch.from_numpy(q['aa']).to(dtype), t=torch.from_numpy(q['t']).to(dtype), K=torch.from_numpy(q['K']).to(dtype),
                noise=torch.from_numpy(q['noise']).to(dtype))


def run_reference(d: dict, cfg: dict) -> dict:
    """The reference's own code path: trainer.py:316-321 (postprocess) + handlers.image_recon + handlers.disp_smooth."""
    ref_shim.load()
    from src import losses, regularizers
    from src.core import handlers
    from src.tools import T_from_AAt, ViewSynth, ops, to_inv, to_scaled

    dt = d['imgs'].dtype
    disps = [x.clone().requires_grad_() for x in d['disps']]
    aa, t, K = (d[k].clone().requires_grad_() for k in ('aa', 't', 'K'))
    H, W = d['imgs'].shape[-2:]
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    to_depth = (lambda x: to_scaled(x, mn, mx)[1]) if (mn or mx) else to_inv
    Ts = T_from_AAt(aa, t)
    depth_up = {s: to_depth(ops.interpolate_like(x, d['imgs'], mode='bilinear')) for s, x in enumerate(disps)}
    crit = losses.ReconstructionLoss(loss_name=cfg.get('loss_name', 'ssim'), use_min=cfg.get('use_min', True),
                                     use_automask=cfg.get('use_automask', True))
    orig = torch.randn_like
    torch.randn_like = lambda x: d['noise']
    try:
        l1, ld1 = handlers.image_recon(crit, ViewSynth((H, W)).to(dt), depth_up, None, d['imgs'], d['supp_imgs'], Ts, K)
    finally:
        torch.randn_like = orig
    l2, ld2 = handlers.disp_smooth(regularizers.SmoothReg(use_edges=True), dict(enumerate(disps)), d['imgs'])
    (l1 + 1e-3*l2).backward()
    out = dict(loss_recon=l1.detach(), loss_smooth=l2.detach(), g_aa=aa.grad, g_t=t.grad, g_K=K.grad,
               warp0=ld1['supp_imgs_warp'].detach().float(), disp_grad=ld2['disp_grad'].detach().float(),
               image_grad=ld2['image_grad'].detach().float(), depth_up0=depth_up[0].detach().float())
    for s, x in enumerate(disps): out[f'g_disp{s}'] = x.grad
    if 'automask' in ld1: out['automask0'] = ld1['automask'].to(torch.uint8)
    return out


def main() -> None:
    GOLDEN.mkdir(parents=True, exist_ok=True)
    for name, cfg in CASES.items():
        q = quantised_inputs(cfg)
        ref64 = run_reference(dequantise(q, torch.float64), cfg)
        ref32 = run_reference(dequantise(q, torch.float32), cfg)
        # The tie-break noise is regenerated in the tests from np.random.RandomState(seed + 100) (stable legacy generator).
        arrays = {f'in_{k}': v for k, v in q.items() if k not in ('disps_u16', 'noise')}
        for s, x in enumerate(q['disps_u16']): arrays[f'in_disp{s}_u16'] = x
        for k, v in ref64.items():
            v = v.numpy()
            if k in ('warp0', 'depth_up0', 'disp_grad', 'image_grad'): v = v[..., ::4, ::4]  # Strided probe keeps the file small.
            elif k.startswith('g_disp'): v = v.astype(np.float32)
            arrays[f'ref64_{k}'] = v
        arrays['ref32_loss_recon'] = ref32['loss_recon'].numpy(); arrays['ref32_loss_smooth'] = ref32['loss_smooth'].numpy()
        for k in ('g_aa', 'g_t', 'g_K'): arrays[f'ref32_{k}'] = ref32[k].numpy()
        arrays['cfg'] = np.array(repr(cfg))
        np.savez_compressed(GOLDEN/f'loss_{name}.npz', **arrays)
        print(f'{name}: recon={ref64["loss_recon"].item():.8f} smooth={ref64["loss_smooth"].item():.8f} '
              f'automask_on={ref64["automask0"].float().mean().item() if "automask0" in ref64 else float("nan"):.3f} '
              f'-> {(GOLDEN/f"loss_{name}.npz").stat().st_size/1024:.0f} KiB')


if __name__ == '__main__':
    main()
