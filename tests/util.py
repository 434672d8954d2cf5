"""Shared helpers of the test-suite: golden fixtures, oracle runner, CUDA runner, error metrics."""
from __future__ import annotations

import ast
from pathlib import Path

import numpy as np
import torch

from oracle import loss as OL

GOLDEN = Path(__file__).resolve().parent/'golden'
LOSS_CASES = ['c1_s1', 'c1_s4', 'ragged_n4', 'mean_noauto_l1', 'behind_ties_const', 'noscale_mean_auto']


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """Norm-wise relative error ||a-b|| / ||b||."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm()/b.norm().clamp(min=1e-30)).item()


def load_golden(name: str):
    """-> (inputs in float64, cfg dict, reference outputs as numpy arrays)."""
    z = np.load(GOLDEN/f'loss_{name}.npz')
    cfg = ast.literal_eval(str(z['cfg']))
    S = cfg['S']
    H, W = cfg['shape']
    noise = np.random.RandomState(cfg['seed'] + 100).standard_normal((S*cfg['b'], 1, H, W)).astype(np.float32)
    f = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    inp = dict(imgs=f(z['in_imgs_u8']).double()/255, supp_imgs=f(z['in_supp_u8']).double()/255,
               disps=[f(z[f'in_disp{s}_u16'].astype(np.int32)).double()/65535 for s in range(S)],
               aa=f(z['in_aa']).double(), t=f(z['in_t']).double(), K=f(z['in_K']).double(), noise=f(noise).double())
    ref = {k: z[k] for k in z.files if k.startswith('ref')}
    return inp, cfg, ref


def cast(inp: dict, dtype, device='cpu') -> dict:
    return {k: ([x.to(device=device, dtype=dtype) for x in v] if isinstance(v, list) else v.to(device=device, dtype=dtype))
            for k, v in inp.items()}


def run_oracle(inp: dict, cfg: dict, dtype=torch.float64, forced_sel=None, w_smooth: float = 1e-3) -> dict:
    d = cast(inp, dtype)
    disps = [x.clone().requires_grad_() for x in d['disps']]
    aa, t, K = (d[k].clone().requires_grad_() for k in ('aa', 't', 'K'))
    Ts = OL.T_from_AAt(aa, t)
    H, W = d['imgs'].shape[-2:]
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    depths = [OL.disp_to_depth(OL.resize_bilinear(x, (H, W)), mn, mx) for x in disps]
    l_rec, o = OL.image_recon(depths, d['imgs'], d['supp_imgs'], Ts, K, cfg.get('use_min', True), cfg.get('use_automask', True),
                              d['noise'], loss_name=cfg.get('loss_name', 'ssim'), forced_sel=forced_sel)
    l_sm, o2 = OL.disp_smooth(disps, d['imgs'], True)
    (l_rec + w_smooth*l_sm).backward()
    out = dict(loss_recon=l_rec.detach(), loss_smooth=l_sm.detach(), g_aa=aa.grad, g_t=t.grad, g_K=K.grad,
               warp0=o['supp_imgs_warp'].detach(), sel=o['sel'], err=o['err'].detach(), depth_up0=depths[0].detach(),
               disp_grad=o2['disp_grad'].detach(), image_grad=o2['image_grad'].detach())
    for s, x in enumerate(disps): out[f'g_disp{s}'] = x.grad
    if 'automask' in o: out['automask0'] = o['automask']
    return out


def run_cuda(inp: dict, cfg: dict, w_smooth: float = 1e-3, device='cuda') -> dict:
    """The product path: libstv kernels through the host-side mirror modules."""
    from slowtv_monodepth_b200 import geometry as G, handlers as Hd
    from slowtv_monodepth_b200.losses import ReconstructionLoss
    from slowtv_monodepth_b200.regularizers import SmoothReg

    d = cast(inp, torch.float32, device)
    disps = [x.clone().requires_grad_() for x in d['disps']]
    aa, t, K = (d[k].clone().requires_grad_() for k in ('aa', 't', 'K'))
    Ts = G.T_from_AAt(aa, t)
    H, W = d['imgs'].shape[-2:]
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    depths = {s: G.upsample_to_depth(x, (H, W), mn, mx)[1] for s, x in enumerate(disps)}
    crit = ReconstructionLoss(cfg.get('loss_name', 'ssim'), cfg.get('use_min', True), cfg.get('use_automask', True))
    l_rec, ld, sel, warp0 = crit.fused(list(depths.values()), d['imgs'], d['supp_imgs'], Ts, K, noise=d['noise'], want_warp=True)
    o = {'supp_imgs_warp': warp0, **{k: v[0] for k, v in ld.items()}}
    l_sm, o2 = Hd.disp_smooth(SmoothReg(use_edges=True), dict(enumerate(disps)), d['imgs'])
    (l_rec + w_smooth*l_sm).backward()
    out = dict(loss_recon=l_rec.detach(), loss_smooth=l_sm.detach(), g_aa=aa.grad, g_t=t.grad, g_K=K.grad,
               warp0=o['supp_imgs_warp'], depth_up0=depths[0].detach(), disp_grad=o2['disp_grad'], image_grad=o2['image_grad'],
               sel=sel.flatten(0, 1).unsqueeze(1))
    for s, x in enumerate(disps): out[f'g_disp{s}'] = x.grad
    if 'automask' in o: out['automask0'] = o['automask']
    return out
