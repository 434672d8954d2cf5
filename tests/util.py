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


class eps32:
    """Context: the oracle uses float32's eps whatever dtype it runs in (the kernels are float32, like the reference's
    training configuration), so a float64 oracle run is the exact-arithmetic version of the float32 reference."""
    def __enter__(self): self.old, OL.FORCE_EPS32 = OL.FORCE_EPS32, True
    def __exit__(self, *a): OL.FORCE_EPS32 = self.old


def run_oracle(inp: dict, cfg: dict, dtype=torch.float64, forced_sel=None, w_smooth: float = 1e-3) -> dict:
    d = cast(inp, dtype)
    disps = [x.clone().requires_grad_() for x in d['disps']]
    aa, t, K = (d[k].clone().requires_grad_() for k in ('aa', 't', 'K'))
    Ts = OL.T_from_AAt(aa, t)
    H, W = d['imgs'].shape[-2:]
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    disps_up = [OL.resize_bilinear(x, (H, W))*1 for x in disps]   # (*1: a node of its own even when the resize is the identity)
    depths = [OL.disp_to_depth(x, mn, mx) for x in disps_up]
    l_rec, o = OL.image_recon(depths, d['imgs'], d['supp_imgs'], Ts, K, cfg.get('use_min', True), cfg.get('use_automask', True),
                              d['noise'], loss_name=cfg.get('loss_name', 'ssim'), forced_sel=forced_sel)
    l_sm, o2 = OL.disp_smooth(disps, d['imgs'], True)
    for x in depths + disps_up: x.retain_grad()
    (l_rec + w_smooth*l_sm).backward()
    out = dict(loss_recon=l_rec.detach(), loss_smooth=l_sm.detach(), g_aa=aa.grad, g_t=t.grad, g_K=K.grad,
               warp0=o['supp_imgs_warp'].detach(), sel=o['sel'], err=o['err'].detach(), depth_up0=depths[0].detach(),
               disp_grad=o2['disp_grad'].detach(), image_grad=o2['image_grad'].detach())
    for s, x in enumerate(disps): out[f'g_disp{s}'] = x.grad
    for s, x in enumerate(depths): out[f'g_depth{s}'] = x.grad
    for s, x in enumerate(disps_up): out[f'g_dispup{s}'] = x.grad   # photometric term only (the smoothness acts on `disps`)
    if 'automask' in o: out['automask0'] = o['automask']
    return out


def oracle_pull_to_disp(inp: dict, cfg: dict, g_full: list, kind: str = 'depth', w_smooth: float = 1e-3) -> list:
    """d loss / d disp_s given the full-resolution gradient maps `g_full[s]` w.r.t. depth_up ('depth') or disp_up ('disp'): the
    vector-Jacobian product of the oracle's bilinear upsample (+ to_scaled / to_inv) plus the oracle's smoothness gradient.
    Both are free of discrete events, so the product's low-resolution gradients can be held to this on EVERY pixel."""
    d = cast(inp, torch.float64)
    H, W = d['imgs'].shape[-2:]
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    disps = [x.clone().requires_grad_() for x in d['disps']]
    l_sm, _ = OL.disp_smooth(disps, d['imgs'], True)
    total = w_smooth*l_sm
    for x, g in zip(disps, g_full):
        up = OL.resize_bilinear(x, (H, W))
        if kind == 'depth': up = OL.disp_to_depth(up, mn, mx)
        total = total + (up*g.detach().double().cpu()).sum()
    total.backward()
    return [x.grad for x in disps]


def unstable_pixels(inp: dict, cfg: dict, tol_val: float = 4e-6, tol_pos: float = 5e-4, pos_err: float = 2e-4):
    """Pixels whose float32 result may legitimately differ O(1) from the exact one because a *discrete* event sits within
    float32 rounding of flipping (computed from the float64 oracle):
      - |warp - target| < tol_val + pos_err * (|d warp/d ix| + |d warp/d iy|) on some channel (sign of the L1 sub-gradient): a
        float32 sample position up to 1024 is only known to ~1e-4 px (ulp(512..1024) = 6e-5, a few roundings in the projection
        chain), so ANY float32 evaluation of the warped value is off by that times the local image gradient,
      - sample position within tol_pos of a texel boundary or of the image border (bilinear gradient is discontinuous),
      - projected depth within tol_val of the 0.1 clamp.
    -> (bad (S*b,1,H,W) bool, candidate errors (S*b, n[+1], H, W) for decision margins)."""
    d = cast(inp, torch.float64)
    n, S = d['supp_imgs'].shape[0], cfg['S']
    H, W = d['imgs'].shape[-2:]
    Ts = OL.T_from_AAt(d['aa'], d['t'])
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    dep = torch.cat([OL.disp_to_depth(OL.resize_bilinear(x, (H, W)), mn, mx) for x in d['disps']], 0)
    tgt = d['imgs'].repeat(S, 1, 1, 1)
    name = cfg.get('loss_name', 'ssim')
    fn = OL.photo_error if name == 'ssim' else (lambda p, t: (p - t).abs().mean(1, keepdim=True))
    bad = torch.zeros(S*cfg['b'], 1, H, W, dtype=torch.bool)
    errs = []
    for k in range(n):
        Tk, Kk = Ts[k].repeat(S, 1, 1), d['K'].repeat(S, 1, 1)
        ix, iy, z, _ = OL.warp_coords(dep, Tk, Kk)
        src = d['supp_imgs'][k].repeat(S, 1, 1, 1)
        w = OL.sample_bilinear_border(src, ix, iy)
        errs.append(fn(w, tgt))
        h = 1e-3   # finite differences of the (piecewise linear) sampler: the steeper side of each axis
        slope = sum(torch.maximum((OL.sample_bilinear_border(src, ix + dx, iy + dy) - w).abs(),
                                  (OL.sample_bilinear_border(src, ix - dx, iy - dy) - w).abs())/h for dx, dy in ((h, 0.), (0., h)))
        bad |= ((w - tgt).abs() < tol_val + pos_err*slope).any(1, keepdim=True)
        for c, size in ((ix, W), (iy, H)):
            cc = c.clamp(0, size - 1)
            bad |= (((cc - cc.round()).abs() < tol_pos) & (c > -tol_pos) & (c < size - 1 + tol_pos)).unsqueeze(1)
        bad |= ((z - 0.1).abs() < tol_val)
    cands = torch.cat(errs, 1)
    if not cfg.get('use_min', True): cands = cands.mean(1, keepdim=True)
    if cfg.get('use_automask', True):
        st = OL.compute_photo(d['supp_imgs'].repeat(1, S, 1, 1, 1), tgt, cfg.get('use_min', True), name)
        cands = torch.cat([cands, st + OL.EPS32*d['noise']], 1)
    return bad, cands


def footprint(bad: torch.Tensor, f: int) -> torch.Tensor:
    """Low-resolution pixels (factor f) whose bilinear-upsampling footprint touches a bad full-resolution pixel."""
    import torch.nn.functional as F
    x = bad.float()
    if f == 1: return bad
    x = F.max_pool2d(x, kernel_size=2*f + 1, stride=1, padding=f)
    return F.max_pool2d(x, kernel_size=f, stride=f) > 0


def check_pixel_gradients(inp: dict, cfg: dict, got: dict, want: dict, bad: torch.Tensor, tol: float, min_frac: float = 0.85,
                          ref32: dict | None = None) -> None:
    """(1) full-resolution maps d loss/d depth_up (or d/d disp_up, `got['up_kind']`) on the stable pixels, which must be at least
    `min_frac` of every scale; (2) low-resolution d loss/d disp_s on every pixel against the oracle's pull-back of those maps.
    The bar of (1) is `tol`, or — where the reference's OWN float32 arithmetic (`ref32`: the oracle run in float32 with the same
    decisions) is further than that from the float64 answer on the same pixels — twice that noise floor."""
    import torch.nn.functional as F
    S, b = cfg['S'], cfg['b']
    kind = got.get('up_kind', 'depth')
    key = 'g_depth' if kind == 'depth' else 'g_dispup'
    for s in range(S):
        good = ~bad[s*b:(s + 1)*b]
        good = ~(F.max_pool2d((~good).float(), 3, 1, 1) > 0)  # a flipped event changes the SSIM/L1 gradient of its 3x3 neighbourhood
        frac = good.float().mean().item()
        assert frac >= min_frac, f'scale {s}: only {frac:.1%} of the pixels are compared'
        e = rel_masked(got[f'g_up{s}'], want[f'{key}{s}'], good)
        floor = rel_masked(ref32[f'{key}{s}'], want[f'{key}{s}'], good) if ref32 is not None else 0.
        assert e < max(tol, 2*floor), f'{key}{s}: {e:.3e} on {frac:.1%} of the pixels (float32 reference noise floor {floor:.3e})'
    pulled = oracle_pull_to_disp(inp, cfg, [got[f'g_up{s}'] for s in range(S)], kind)
    for s in range(S):
        e = rel(got[f'g_disp{s}'], pulled[s])
        assert e < 0.1*tol, f'g_disp{s} vs the oracle pull-back of the full-resolution map: {e:.3e}'


def rel_masked(a, b, good) -> float:
    a, b = a.detach().double().cpu()*good, b.detach().double().cpu()*good
    return ((a - b).norm()/b.norm().clamp(min=1e-30)).item()


def run_cuda(inp: dict, cfg: dict, w_smooth: float = 1e-3, device='cuda', mode: str = 'fused-disp') -> dict:
    """The product path: libstv kernels through the host-side mirror modules.
      mode 'fused-disp'   single-pass kernel fed with the low-resolution disparities (what the training step runs);
      mode 'fused-depth'  single-pass kernel fed with up-sampled depth maps (stv_disp_to_depth_fwd/bwd around it: what a caller of
                          the reference's handlers.image_recon signature gets);
      mode 'two-pass'     stv_photo_fwd + stv_photo_bwd (always used for the mean reduction)."""
    from slowtv_monodepth_b200 import functional as F_, geometry as G, handlers as Hd
    from slowtv_monodepth_b200.losses import ReconstructionLoss
    from slowtv_monodepth_b200.regularizers import SmoothReg

    d = cast(inp, torch.float32, device)
    disps = [x.clone().requires_grad_() for x in d['disps']]
    aa, t, K = (d[k].clone().requires_grad_() for k in ('aa', 't', 'K'))
    Ts = G.T_from_AAt(aa, t)
    H, W = d['imgs'].shape[-2:]
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    use_min = cfg.get('use_min', True)
    crit = ReconstructionLoss(cfg.get('loss_name', 'ssim'), use_min, cfg.get('use_automask', True))
    F_.PHOTO_FORCE_TWO_PASS, F_.KEEP_UNIT_GRADS = mode == 'two-pass', True
    try:
        if mode == 'fused-disp' and use_min:
            l_rec, ld, sel, warp0 = crit.fused(disps, d['imgs'], d['supp_imgs'], Ts, K, noise=d['noise'], want_warp=True, from_disp=(mn, mx))
            depths, kind = None, 'disp'
        else:
            depths = {s: G.upsample_to_depth(x, (H, W), mn, mx)[1] for s, x in enumerate(disps)}
            for x in depths.values(): x.retain_grad()
            l_rec, ld, sel, warp0 = crit.fused(list(depths.values()), d['imgs'], d['supp_imgs'], Ts, K, noise=d['noise'], want_warp=True)
            kind = 'depth'
        unit = F_.LAST_UNIT_GRADS
    finally:
        F_.PHOTO_FORCE_TWO_PASS, F_.KEEP_UNIT_GRADS, F_.LAST_UNIT_GRADS = False, False, None
    o = {'supp_imgs_warp': warp0, **{k: v[0] for k, v in ld.items()}}
    l_sm, o2 = Hd.disp_smooth(SmoothReg(use_edges=True), dict(enumerate(disps)), d['imgs'])
    (l_rec + w_smooth*l_sm).backward()
    with torch.no_grad(): depth_up0 = G.upsample_to_depth(d['disps'][0], (H, W), mn, mx)[1]
    out = dict(loss_recon=l_rec.detach(), loss_smooth=l_sm.detach(), g_aa=aa.grad, g_t=t.grad, g_K=K.grad,
               warp0=o['supp_imgs_warp'], depth_up0=depth_up0, disp_grad=o2['disp_grad'], image_grad=o2['image_grad'],
               sel=sel.flatten(0, 1).unsqueeze(1), up_kind=kind)
    for s, x in enumerate(disps): out[f'g_disp{s}'] = x.grad
    for s in range(len(disps)): out[f'g_up{s}'] = depths[s].grad if depths is not None else unit[s]  # dL/dloss_recon = 1 here
    if 'automask' in o: out['automask0'] = o['automask']
    return out

def check_loss_stack(inp: dict, cfg: dict, got: dict, tol_loss: float = 1e-5, tol_grad: float = 1e-4) -> None:
    """The parity protocol of tests/test_loss_gpu.py (see its docstring) for one set of inputs and one CUDA result."""
    import pytest
    sel = got['sel'].cpu()
    S, b = cfg['S'], cfg['b']
    use_min, use_auto = cfg.get('use_min', True), cfg.get('use_automask', True)

    with eps32():
        bad, cands = unstable_pixels(inp, cfg)
        # (1) decisions
        free = run_oracle(inp, cfg, torch.float64)
        if use_min or use_auto:
            osel = free['sel']
            if not use_min: osel = torch.where(osel == 255, osel, torch.full_like(osel, 254))
            diff = sel != osel
            top2 = cands.topk(2, dim=1, largest=False)[0]
            margin = top2[:, 1:2] - top2[:, 0:1]
            assert diff.float().mean().item() < 5e-3, f'{diff.float().mean().item():.4%} decisions differ'
            assert (margin[diff] < 2e-5).all(), f'decision flipped with margin {margin[diff].max().item():.3e}'

        # (2) values and gradients given the kernel's decisions
        if not use_min and use_auto:
            pytest.skip('mean-reduction + automask has no forced-decision mode in the oracle; covered by (1) and the golden test')
        want = run_oracle(inp, cfg, torch.float64, forced_sel=sel if use_min else None)

    assert bad.float().mean().item() < 0.05, f'{bad.float().mean().item():.2%} unstable pixels'
    assert rel(got['loss_recon'], want['loss_recon']) < tol_loss
    assert rel(got['loss_smooth'], want['loss_smooth']) < tol_loss
    # Pose / intrinsics gradients are sums over ALL pixels, including those within float32 rounding of a sub-gradient event (whose
    # per-pixel term legitimately differs O(1) between any two float32 evaluations). The bar is tol_grad, or — where the
    # reference's OWN float32 arithmetic (the oracle run in float32 with the same decisions) is further than that from the
    # float64 answer — twice that noise floor.
    with eps32():
        ref32 = run_oracle(inp, cfg, torch.float32, forced_sel=sel if use_min else None)
    for k in ('g_aa', 'g_t', 'g_K'):
        floor = rel(ref32[k], want[k])
        assert rel(got[k], want[k]) < max(tol_grad, 2*floor), f'{k}: {rel(got[k], want[k]):.3e} (float32 reference noise floor {floor:.3e})'
    check_pixel_gradients(inp, cfg, got, want, bad, tol_grad, ref32=ref32)
    for k in ('warp0', 'depth_up0', 'disp_grad', 'image_grad'):
        assert rel(got[k], want[k]) < 1e-5, f'{k}: {rel(got[k], want[k]):.3e}'
