"""ORACLE (test infrastructure — not shipped, not on the product path).

CPU restatement of the reference's view-synthesis loss stack, written as explicit per-pixel arithmetic with
plain PyTorch CPU tensors (dtype-generic: run it in float64 for the ground truth, float32 for a like-for-like
check). Gradients come from autograd over this explicit arithmetic. It deliberately does NOT call
`F.grid_sample`, `F.interpolate`, `nn.AvgPool2d` or `nn.ReflectionPad2d`: those are what the reference calls, and
`tests/test_oracle_vs_reference.py` + `oracle/make_golden.py` pin this restatement against the reference's own
code executed in the build container (fixtures under `tests/golden/`).

Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / reference arm of `bench.py` may import this.

Every function cites the reference file:line (under /root/reference) that it restates.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

EPS32 = 1.1920928955078125e-07  # torch.finfo(torch.float32).eps; reference `ops.eps` (src/tools/ops.py:63-66)


FORCE_EPS32 = False  # Tests set this so a float64 run models "the float32 reference with exact arithmetic".


def _eps(x: Tensor) -> float:
    return EPS32 if FORCE_EPS32 else torch.finfo(x.dtype).eps


# ---------------------------------------------------------------------------------------------------------------------
# Row 6: bilinear upsample, align_corners=False  (src/tools/ops.py:311-314 -> F.interpolate; used trainer.py:320)
# ---------------------------------------------------------------------------------------------------------------------
def _lin_taps(n_out: int, n_in: int, dtype, device):
    """Source taps for ATen `upsample_bilinear2d(align_corners=False)` along one axis.
    src = max(scale*(dst+0.5)-0.5, 0), i0=floor(src), i1=min(i0+1, n_in-1), lam=src-i0."""
    scale = n_in/n_out
    dst = torch.arange(n_out, dtype=dtype, device=device)
    src = (scale*(dst + 0.5) - 0.5).clamp(min=0)
    i0 = src.floor().long().clamp(max=n_in-1)
    i1 = (i0 + 1).clamp(max=n_in-1)
    lam = src - i0.to(dtype)
    return i0, i1, lam


def resize_bilinear(x: Tensor, size: tuple[int, int]) -> Tensor:
    """(b, c, h, w) -> (b, c, H, W). Used both to upsample disparities (trainer.py:320) and to downsample
    the target image to each disparity scale (handlers.py:278)."""
    H, W = size
    h, w = x.shape[-2:]
    if (h, w) == (H, W): return x
    y0, y1, ly = _lin_taps(H, h, x.dtype, x.device)
    x0, x1, lx = _lin_taps(W, w, x.dtype, x.device)
    ly = ly[:, None]
    top = x[..., y0, :]
    bot = x[..., y1, :]
    rows = top*(1 - ly) + bot*ly  # (b, c, H, w)
    out = rows[..., x0]*(1 - lx) + rows[..., x1]*lx
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Row 7: disparity -> depth  (src/tools/geometry.py:62-76, 86-90; trainer.py:49,321)
# ---------------------------------------------------------------------------------------------------------------------
def to_inv(depth: Tensor) -> Tensor:
    return (depth > 0).to(depth.dtype) / depth.clamp(min=_eps(depth))


def disp_to_depth(disp: Tensor, min_depth: float | None, max_depth: float | None) -> Tensor:
    if not (min_depth or max_depth): return to_inv(disp)
    i_max, i_min = 1/min_depth, (1/max_depth) if max_depth else 0
    return to_inv((i_max - i_min)*disp + i_min)


# ---------------------------------------------------------------------------------------------------------------------
# Row 5: pose / intrinsics helpers  (src/tools/geometry.py:136-140, 181-209, 249-263; src/networks/pose.py:61-73)
# ---------------------------------------------------------------------------------------------------------------------
def T_from_AAt(aa: Tensor, t: Tensor) -> Tensor:
    """Rodrigues: T = I + W sin(a) + W^2 (1 - cos(a)), axis = aa/max(|aa|, eps), T[:3,3] = t."""
    angle = aa.norm(p=2, dim=-1, keepdim=True)
    axis = aa/angle.clamp(min=_eps(aa))
    x, y, z = axis.unbind(-1)
    zr = torch.zeros_like(x)
    Wm = torch.stack([
        torch.stack([zr, -z, y], -1),
        torch.stack([z, zr, -x], -1),
        torch.stack([-y, x, zr], -1)], -2)
    a = angle.unsqueeze(-1)
    R = torch.eye(3, dtype=aa.dtype, device=aa.device) + Wm*a.sin() + (Wm @ Wm)*(1 - a.cos())
    T = torch.zeros(aa.shape[:-1] + (4, 4), dtype=aa.dtype, device=aa.device)
    T = T.clone()
    top = torch.cat([R, t.unsqueeze(-1)], -1)
    bot = torch.tensor([0, 0, 0, 1], dtype=aa.dtype, device=aa.device).expand(aa.shape[:-1] + (1, 4))
    return torch.cat([top, bot], -2)


def build_K(fs: Tensor, cs: Tensor) -> Tensor:
    b = fs.shape[0]
    K = torch.eye(4, dtype=fs.dtype, device=fs.device).repeat(b, 1, 1)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = fs[:, 0], fs[:, 1], cs[:, 0], cs[:, 1]
    return K


def resize_K(K: Tensor, new_shape: tuple[int, int], shape: tuple[int, int] = (1, 1)) -> Tensor:
    K = K.clone()
    K[..., 0, :] = K[..., 0, :]*(new_shape[1]/shape[1])
    K[..., 1, :] = K[..., 1, :]*(new_shape[0]/shape[0])
    return K


# ---------------------------------------------------------------------------------------------------------------------
# Rows 9-11: backproject -> rigid transform -> project -> bilinear sample  (src/tools/geometry.py:285-391)
# ---------------------------------------------------------------------------------------------------------------------
def warp_coords(depth: Tensor, T: Tensor, K: Tensor, K_inv: Tensor | None = None):
    """depth (B,1,H,W), T/K (B,4,4) -> sample coords in *pixel units* (ix, iy) each (B,H,W), and the
    depth in the new frame (B,1,H,W).

    P = depth * Kinv3x3 (u,v,1)                       geometry.py:313-315
    Q = R P + t                                       geometry.py:386
    zc = max(max(Qz, eps), 0.1);  q = K3x3 (Q/zc)     geometry.py:339-341
    grid = (q/(size-1) - .5)*2                        geometry.py:347-349
    pixel = ((grid+1)*size - 1)/2                     ATen grid_sampler_unnormalize(align_corners=False)
    """
    B, _, H, W = depth.shape
    dt, dev = depth.dtype, depth.device
    if K_inv is None: K_inv = torch.linalg.inv(K)
    u = torch.arange(W, dtype=dt, device=dev).view(1, 1, W).expand(B, H, W)
    v = torch.arange(H, dtype=dt, device=dev).view(1, H, 1).expand(B, H, W)
    Ki = K_inv[:, :3, :3]
    e = lambda M, i, j: M[:, i, j].view(B, 1, 1)
    d = depth[:, 0]
    ray = [e(Ki, i, 0)*u + e(Ki, i, 1)*v + e(Ki, i, 2) for i in range(3)]
    P = [r*d for r in ray]
    Q = [e(T, i, 0)*P[0] + e(T, i, 1)*P[1] + e(T, i, 2)*P[2] + e(T, i, 3) for i in range(3)]
    z = Q[2].clamp(min=_eps(depth))
    zc = z.clamp(min=0.1)
    n = [q/zc for q in Q]
    qx = e(K, 0, 0)*n[0] + e(K, 0, 1)*n[1] + e(K, 0, 2)*n[2]
    qy = e(K, 1, 0)*n[0] + e(K, 1, 1)*n[1] + e(K, 1, 2)*n[2]
    gx = (qx/(W - 1) - 0.5)*2
    gy = (qy/(H - 1) - 0.5)*2
    ix = ((gx + 1)*W - 1)/2
    iy = ((gy + 1)*H - 1)/2
    return ix, iy, z.unsqueeze(1), (gx, gy)


def _clip_border(c: Tensor, size: int) -> Tensor:
    """ATen `clip_coordinates_set_grad`: clamp to [0, size-1]; the borders themselves count as out of bounds for the
    gradient (zero gradient when c <= 0 or c >= size-1)."""
    inside = (c > 0) & (c < size - 1)
    return torch.where(inside, c, c.detach().clamp(0, size - 1))


def sample_bilinear_border(img: Tensor, ix: Tensor, iy: Tensor) -> Tensor:
    """F.grid_sample(mode='bilinear', padding_mode='border', align_corners=False) given pixel-unit coords.
    img (B,C,H,W); ix, iy (B,H',W') -> (B,C,H',W').  (geometry.py:364,389)"""
    B, C, H, W = img.shape
    ix, iy = _clip_border(ix, W), _clip_border(iy, H)
    x0f, y0f = ix.floor(), iy.floor()
    wx1, wy1 = ix - x0f, iy - y0f
    wx0, wy0 = 1 - wx1, 1 - wy1
    x0, y0 = x0f.long(), y0f.long()
    x1, y1 = x0 + 1, y0 + 1
    flat = img.reshape(B, C, H*W)

    def tap(yy, xx):
        ok = ((xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)).to(img.dtype)  # Taps outside the image add nothing.
        idx = (yy.clamp(0, H - 1)*W + xx.clamp(0, W - 1)).reshape(B, 1, -1).expand(B, C, -1)
        return flat.gather(2, idx).reshape(B, C, *ix.shape[1:]) * ok.unsqueeze(1)

    out = (tap(y0, x0)*(wy0*wx0).unsqueeze(1) + tap(y0, x1)*(wy0*wx1).unsqueeze(1) +
           tap(y1, x0)*(wy1*wx0).unsqueeze(1) + tap(y1, x1)*(wy1*wx1).unsqueeze(1))
    return out


def view_synth(inp: Tensor, depth: Tensor, T: Tensor, K: Tensor, K_inv: Tensor | None = None):
    """`ViewSynth.forward` (geometry.py:366-391) -> (input_warp, depth_warp, mask_valid)."""
    ix, iy, dw, (gx, gy) = warp_coords(depth, T, K, K_inv)
    mask_valid = ((gx.abs() < 1) & (gy.abs() < 1)).unsqueeze(1)
    return sample_bilinear_border(inp, ix, iy), dw, mask_valid


# ---------------------------------------------------------------------------------------------------------------------
# Rows 12-13: SSIM + L1 photometric error  (src/losses/photometric.py:11-14, 23-51, 54-88)
# ---------------------------------------------------------------------------------------------------------------------
def _reflect_pad1(x: Tensor) -> Tensor:
    """nn.ReflectionPad2d(1): index -1 -> 1, n -> n-2."""
    x = torch.cat([x[..., 1:2, :], x, x[..., -2:-1, :]], -2)
    x = torch.cat([x[..., :, 1:2], x, x[..., :, -2:-1]], -1)
    return x


def _box3(x: Tensor) -> Tensor:
    """nn.AvgPool2d(3, 1) on an already padded map."""
    H, W = x.shape[-2] - 2, x.shape[-1] - 2
    acc = 0
    for dy in range(3):
        for dx in range(3):
            acc = acc + x[..., dy:dy + H, dx:dx + W]
    return acc/9


def ssim_error(pred: Tensor, target: Tensor) -> Tensor:
    C1, C2 = 0.01**2, 0.03**2
    x, y = _reflect_pad1(pred), _reflect_pad1(target)
    mu_x, mu_y = _box3(x), _box3(y)
    sig_x = _box3(x*x) - mu_x*mu_x
    sig_y = _box3(y*y) - mu_y*mu_y
    sig_xy = _box3(x*y) - mu_x*mu_y
    num = (2*mu_x*mu_y + C1)*(2*sig_xy + C2)
    den = (mu_x*mu_x + mu_y*mu_y + C1)*(sig_x + sig_y + C2)
    return ((1 - num/den)/2).clamp(0, 1)


def photo_error(pred: Tensor, target: Tensor, weight_ssim: float = 0.85) -> Tensor:
    """(b,3,h,w) x2 -> (b,1,h,w)."""
    return (weight_ssim*ssim_error(pred, target).mean(1, keepdim=True) +
            (1 - weight_ssim)*(pred - target).abs().mean(1, keepdim=True))


# ---------------------------------------------------------------------------------------------------------------------
# Row 14: min-reprojection / automask  (src/losses/reconstruction.py:43-44, 59-77, 79-96, 98-126)
# ---------------------------------------------------------------------------------------------------------------------
def compute_photo(pred: Tensor, target: Tensor, use_min: bool = True, loss_name: str = 'ssim') -> Tensor:
    """pred (n,b,3,h,w), target (b,3,h,w) -> (b,1,h,w)."""
    fn = photo_error if loss_name == 'ssim' else (lambda p, t: (p - t).abs().mean(1, keepdim=True))
    err = torch.stack([fn(p, target)[:, 0] for p in pred], 1)  # (b, n, h, w)
    return err.min(1, keepdim=True)[0] if use_min else err.mean(1, keepdim=True)


def reconstruction_loss(pred: Tensor, target: Tensor, source: Tensor | None, use_min: bool = True,
                        use_automask: bool = True, noise: Tensor | None = None, loss_name: str = 'ssim',
                        forced_sel: Tensor | None = None):
    """-> (loss, automask|None, err_map, sel).

    `noise` replaces `torch.randn_like(err_static)` (reconstruction.py:72) so that both sides of a parity check see the
    same tie-break noise; `None` means zero noise.

    `sel` (b,1,h,w) uint8 is the discrete decision taken per pixel: k in [0,n) = warped support k carries the loss,
    255 = the static (identity) error won, i.e. the pixel is auto-masked. Passing `forced_sel` replaces the argmin by the
    given decisions: value and gradient are then smooth in the inputs, which is how the CUDA path is compared
    at tight tolerance (decisions themselves are compared separately, modulo near-ties). Requires `use_min`."""
    fn = photo_error if loss_name == 'ssim' else (lambda p, t: (p - t).abs().mean(1, keepdim=True))
    errs = torch.stack([fn(p, target)[:, 0] for p in pred], 1)  # (b, n, h, w)
    if use_min: err, sel = errs.min(1, keepdim=True)  # First index wins ties.
    else: err, sel = errs.mean(1, keepdim=True), torch.zeros_like(errs[:, :1], dtype=torch.long)
    automask, err_static = None, None
    if use_automask:
        err_static = compute_photo(source, target, use_min, loss_name)
        if noise is not None: err_static = err_static + _eps(err_static)*noise
        both = torch.cat([err, err_static], 1)
        err, idx = both.min(1, keepdim=True)  # First index wins ties.
        automask = idx == 0
        sel = torch.where(automask, sel, torch.full_like(sel, 255))
    if forced_sel is not None:
        assert use_min, 'forced decisions are only defined for min-reprojection'
        fs = forced_sel.long()
        static = fs == 255
        err_w = errs.gather(1, fs.clamp(max=errs.shape[1] - 1))
        err = torch.where(static, err_static if err_static is not None else err_w, err_w)
        automask, sel = ~static, fs
    return err.mean(), automask, err, sel.to(torch.uint8)


def dense_error(pred: Tensor, target: Tensor, loss_name: str = 'ssim') -> Tensor:
    """(b,c,h,w) x2 -> (b,1,h,w): PhotoError(0.85) | DenseL1Error | DenseL2Error (src/losses/photometric.py:12-23, 54-88)."""
    if loss_name == 'ssim': return photo_error(pred, target)
    if loss_name == 'l1': return (pred - target).abs().mean(1, keepdim=True)
    if loss_name == 'l2': return (pred - target).pow(2).sum(1, keepdim=True).clamp(min=_eps(pred)).sqrt()
    raise KeyError(loss_name)


def apply_mask(err: Tensor, mask: Tensor | None, mask_name: str | None) -> Tensor:
    """src/losses/reconstruction.py:46-57. err, mask (b,n,h,w)."""
    if mask_name and mask is None: raise ValueError("Must provide a 'mask' when masking...")
    if mask_name == 'explainability': return err*mask
    if mask_name == 'uncertainty': return err*(-mask).exp() + mask
    return err


def reconstruction_loss_ex(pred: Tensor, target: Tensor, source: Tensor | None = None, mask: Tensor | None = None,
                           loss_name: str = 'ssim', use_min: bool = False, use_automask: bool = False, mask_name: str | None = None,
                           noise: Tensor | None = None):
    """The whole registered class (src/losses/reconstruction.py:13-126) on pre-warped frames: pred, source (n,b,c,h,w), target
    (b,c,h,w), mask (b,n,h,w) -> (loss, automask|None, err (b,1,h,w), sel (b,1,h,w) uint8 in the encoding of stv_recon_ex_fwd:
    k | 0x40 (mean); bit 7 = automasked, low bits = the static frame that was the minimum)."""
    def photo(frames):
        e = torch.stack([dense_error(f, target, loss_name)[:, 0] for f in frames], 1)   # (b, n, h, w)
        e = apply_mask(e, mask, mask_name)
        if use_min:
            v, k = e.min(1, keepdim=True)   # first index wins ties
            return v, k
        return e.mean(1, keepdim=True), torch.full_like(e[:, :1], 0x40, dtype=torch.long)
    err, sel = photo(pred)
    automask = None
    if use_automask:
        if source is None: raise ValueError("Must provide the original 'source' images when automasking...")
        es, ksel = photo(source)
        if noise is not None: es = es + _eps(es)*noise
        both = torch.cat([err, es], 1)
        err, idx = both.min(1, keepdim=True)
        automask = idx == 0
        sel = torch.where(automask, sel, ksel | 0x80)
    return err.mean(), automask, err, sel.to(torch.uint8)


def feat_recon(depth0: Tensor, feats: Tensor, supp_feats: Tensor, Ts: Tensor, Ks: Tensor, loss_name: str = 'l2', use_min: bool = False,
               use_automask: bool = False, mask0: Tensor | None = None, mask_name: str | None = None, noise: Tensor | None = None):
    """src/core/handlers.py:70-119: feats (b,c,h4,w4), supp_feats (n,b,c,h4,w4) at 1/4 resolution, depth0 (b,1,H,W).
    -> (loss, warped features (n,b,c,H,W))."""
    size = depth0.shape[-2:]
    f = resize_bilinear(feats.detach(), size)
    sf = torch.stack([resize_bilinear(x.detach(), size) for x in supp_feats])
    warp = torch.stack([view_synth(sf[k], depth0, Ts[k], Ks)[0] for k in range(sf.shape[0])])
    loss = reconstruction_loss_ex(warp, f, sf, mask0, loss_name, use_min, use_automask, mask_name, noise)[0]
    return loss, warp


def regression_loss(pred: Tensor, target: Tensor, mask: Tensor | None = None, loss_name: str = 'berhu', invert: bool = False):
    """RegressionLoss.forward (src/losses/regression.py:11-38, 67-75) -> (loss, masked error map)."""
    if invert: pred, target = to_inv(pred), to_inv(target)
    if mask is None: mask = torch.ones_like(target)
    diff = (pred - target).abs()
    if loss_name == 'l1': e = diff
    elif loss_name == 'log_l1': e = (1 + diff).log()
    elif loss_name == 'berhu':
        delta = 0.2*diff.max()
        e = torch.where(diff <= delta, diff, (diff.pow(2) + delta.pow(2))/(2*delta + _eps(pred)))
    else: raise KeyError(loss_name)
    err = mask*e
    return err.sum()/mask.sum(), err


def stereo_const(disps: list[Tensor], depths: list[Tensor], disps_stereo: list[Tensor], depths_stereo: list[Tensor], T_stereo: Tensor,
                 K: Tensor, loss_name: str = 'l1'):
    """src/core/handlers.py:151-198 -> (loss, warped disparities (2*S*b,1,h,w): virtual-stereo half first)."""
    S = len(disps)
    d, z, ds, zs = (torch.cat(v) for v in (disps, depths, disps_stereo, depths_stereo))   # (S*b, 1, h, w), scale-major
    T = torch.cat([T_stereo]*S)
    all_disps = torch.cat((ds, d))
    warp = view_synth(all_disps, torch.cat((z, zs)), torch.cat((T, torch.linalg.inv(T))), torch.cat([K]*(2*S)))[0]
    return regression_loss(all_disps, warp, None, loss_name)[0], warp


def depth_regr(depths: list[Tensor], targets: Tensor, imgs: Tensor, supp_imgs: Tensor, Ts: Tensor, Ks: Tensor, loss_name: str = 'log_l1',
               invert: bool = False, use_automask: bool = True, photo_name: str = 'ssim', photo_min: bool = True):
    """src/core/handlers.py:201-259 -> (loss, mask (S*b,1,h,w) bool)."""
    S, n = len(depths), supp_imgs.shape[0]
    z, tg, im = torch.cat(depths), torch.cat([targets]*S), torch.cat([imgs]*S)
    masks = tg > 0
    if use_automask:
        def warp_all(dep):
            return torch.stack([view_synth(torch.cat([supp_imgs[k]]*S), dep, torch.cat([Ts[k]]*S), torch.cat([Ks]*S))[0] for k in range(n)])
        automask = compute_photo(warp_all(z), im, photo_min, photo_name) > compute_photo(warp_all(tg), im, photo_min, photo_name)
        masks = masks & automask
    return regression_loss(z, tg, masks.to(z.dtype), loss_name, invert)[0], masks


# ---------------------------------------------------------------------------------------------------------------------
# Row 8: handlers.image_recon  (src/core/handlers.py:14-67)
# ---------------------------------------------------------------------------------------------------------------------
def image_recon(depths: list[Tensor], imgs: Tensor, supp_imgs: Tensor, Ts: Tensor, Ks: Tensor,
                use_min: bool = True, use_automask: bool = True, noise: Tensor | None = None,
                K_inv: Tensor | None = None, loss_name: str = 'ssim', forced_sel: Tensor | None = None):
    """depths: S x (b,1,H,W); imgs (b,3,H,W); supp_imgs (n,b,3,H,W); Ts (n,b,4,4); Ks (b,4,4);
    noise (S*b,1,H,W)|None -> (loss, {'supp_imgs_warp': (n,b,3,H,W), 'automask': (b,1,H,W), 'err': (S*b,1,H,W)})."""
    n, S, b = supp_imgs.shape[0], len(depths), imgs.shape[0]
    dep = torch.cat(depths, 0)  # (S*b,1,H,W), index s*b + i
    tgt = imgs.repeat(S, 1, 1, 1)
    warps, srcs = [], []
    for k in range(n):
        src = supp_imgs[k].repeat(S, 1, 1, 1)
        w, _, _ = view_synth(src, dep, Ts[k].repeat(S, 1, 1), Ks.repeat(S, 1, 1),
                             None if K_inv is None else K_inv.repeat(S, 1, 1))
        warps.append(w); srcs.append(src)
    warps, srcs = torch.stack(warps), torch.stack(srcs)  # (n, S*b, 3, H, W)
    loss, automask, err, sel = reconstruction_loss(warps, tgt, srcs, use_min, use_automask, noise, loss_name, forced_sel)
    out = {'supp_imgs_warp': warps[:, :b], 'err': err, 'sel': sel}
    if automask is not None:
        out['automask'] = automask[:b]
        out['automask_all'] = automask
    return loss, out


# ---------------------------------------------------------------------------------------------------------------------
# Rows 15-16: edge-aware smoothness  (src/regularizers/smooth.py:12-30, 71-97; src/tools/ops.py:279-286;
#                                     src/core/handlers.py:262-281)
# ---------------------------------------------------------------------------------------------------------------------
def _abs_grad(x: Tensor, ch_mean: bool = False):
    dx = torch.zeros_like(x); dy = torch.zeros_like(x)
    dx = torch.cat([(x[..., :, :-1] - x[..., :, 1:]).abs(), torch.zeros_like(x[..., :, :1])], -1)
    dy = torch.cat([(x[..., :-1, :] - x[..., 1:, :]).abs(), torch.zeros_like(x[..., :1, :])], -2)
    if ch_mean: dx, dy = dx.mean(1, keepdim=True), dy.mean(1, keepdim=True)
    return dx, dy


def gaussian_blur3(x: Tensor) -> Tensor:
    """kornia.filters.gaussian_blur2d(x, (3, 3), (1, 1)) (kornia 0.6.10, third-party, absent here — PARITY UNPINNED): separable
    normalised Gaussian taps exp(-d^2/2), d in {-1, 0, 1}, border_type='reflect' (no edge repeat)."""
    g = torch.tensor([-1., 0., 1.], dtype=x.dtype, device=x.device).pow(2).mul(-0.5).exp()
    g = g/g.sum()
    xp = _reflect_pad1(x)
    h = g[0]*xp[..., :, :-2] + g[1]*xp[..., :, 1:-1] + g[2]*xp[..., :, 2:]
    return g[0]*h[..., :-2, :] + g[1]*h[..., 1:-1, :] + g[2]*h[..., 2:, :]


def _abs_grad_ex(x: Tensor, use_blur: bool = False, ch_mean: bool = False):
    """compute_grad (src/regularizers/smooth.py:12-30)."""
    if use_blur: x = gaussian_blur3(x)
    return _abs_grad(x, ch_mean)


def _abs_laplacian(x: Tensor, use_blur: bool = False, ch_mean: bool = False):
    """compute_laplacian (src/regularizers/smooth.py:33-48) -> (dxx, dyy, dxy, dyx)."""
    dx, dy = _abs_grad_ex(x, use_blur)
    dxx, dxy = _abs_grad_ex(dx, use_blur)
    dyx, dyy = _abs_grad_ex(dy, use_blur)
    if ch_mean: dxx, dxy, dyx, dyy = (v.mean(1, keepdim=True) for v in (dxx, dxy, dyx, dyy))
    return dxx, dyy, dxy, dyx


def smooth_reg_ex(disp: Tensor, img: Tensor, use_edges: bool = False, use_laplacian: bool = False, use_blur: bool = False):
    """SmoothReg.forward with every constructor flag (src/regularizers/smooth.py:51-97) -> (loss, disp_grad, image_grad)."""
    eps = _eps(disp)
    fn = _abs_laplacian if use_laplacian else _abs_grad_ex
    d = disp/disp.mean((2, 3), keepdim=True).clamp(min=eps)
    ddx, ddy = fn(d, use_blur)[:2]
    disp_grad = (ddx**2 + ddy**2).clamp(min=eps).sqrt()
    idx, idy = fn(img, use_blur, True)[:2]
    img_grad = (idx**2 + idy**2).clamp(min=eps).sqrt()
    if use_edges: ddx, ddy = ddx*(-idx).exp(), ddy*(-idy).exp()
    return ddx.mean() + ddy.mean(), disp_grad, img_grad


def feat_peak_reg(feat: Tensor, img: Tensor, use_edges: bool = False):
    """FeatPeakReg.forward (src/regularizers/smooth.py:111-136) -> (loss, feat_grad)."""
    fdx, fdy = _abs_grad(feat)
    fg = (fdx**2 + fdy**2).clamp(min=_eps(feat)).sqrt()
    if use_edges:
        dx, dy = _abs_grad(img, ch_mean=True)
        fdx, fdy = fdx*(-dx).exp(), fdy*(-dy).exp()
    return -(fdx.mean() + fdy.mean()), fg


def feat_smooth_reg(feat: Tensor, img: Tensor, use_edges: bool = False):
    """FeatSmoothReg.forward (src/regularizers/smooth.py:150-176) -> (loss, feat_grad)."""
    fxx, fyy, fxy, fyx = _abs_laplacian(feat)
    fg = (fxx**2 + fyy**2).clamp(min=_eps(feat)).sqrt()
    if use_edges:
        dxx, dyy, dxy, dyx = _abs_laplacian(img, ch_mean=True)
        fxx, fyy, fxy, fyx = fxx*(-dxx).exp(), fyy*(-dyy).exp(), fxy*(-dxy).exp(), fyx*(-dyx).exp()
    return fxx.mean() + fyy.mean() + fxy.mean() + fyx.mean(), fg


def mask_reg(x: Tensor) -> Tensor:
    """MaskReg.forward (src/regularizers/mask.py:20-30): F.binary_cross_entropy(x, 1)."""
    return -(x.log().clamp(min=-100)).mean()


def occ_reg(x: Tensor, invert: bool = False) -> Tensor:
    """OccReg.forward (src/regularizers/occlusion.py:31-40)."""
    return (-1 if invert else 1)*x.mean()


def smooth_reg(disp: Tensor, img: Tensor, use_edges: bool = True):
    """-> (loss, disp_grad, image_grad)."""
    eps = _eps(disp)
    d = disp/disp.mean((2, 3), keepdim=True).clamp(min=eps)
    ddx, ddy = _abs_grad(d)
    disp_grad = (ddx**2 + ddy**2).clamp(min=eps).sqrt()
    idx, idy = _abs_grad(img, ch_mean=True)
    img_grad = (idx**2 + idy**2).clamp(min=eps).sqrt()
    if use_edges: ddx, ddy = ddx*(-idx).exp(), ddy*(-idy).exp()
    return ddx.mean() + ddy.mean(), disp_grad, img_grad


def disp_smooth(disps: list[Tensor], imgs: Tensor, use_edges: bool = True, scales: list[int] | None = None):
    """disps: list over scales of (b,1,h_s,w_s); loss = mean_s(loss_s / 2**s)."""
    scales = list(range(len(disps))) if scales is None else scales
    outs = [smooth_reg(d, resize_bilinear(imgs, d.shape[-2:]), use_edges) for d in disps]
    loss = torch.stack([o[0]/2**s for s, o in zip(scales, outs)]).mean()
    return loss, {'disp_grad': outs[0][1], 'image_grad': outs[0][2]}


# ---------------------------------------------------------------------------------------------------------------------
# Rows 1 (loss part): MonoDepthModule.forward_postprocess + forward_loss for the KBR loss set
# (src/core/trainer.py:316-321, 347, 385-393, 436-438, 462)
# ---------------------------------------------------------------------------------------------------------------------
def loss_stack(disps: list[Tensor], imgs: Tensor, supp_imgs: Tensor, Ts: Tensor, Ks: Tensor,
               min_depth: float | None = 0.1, max_depth: float | None = 100., w_recon: float = 1., w_smooth: float = 1e-3,
               use_min: bool = True, use_automask: bool = True, use_edges: bool = True, noise: Tensor | None = None,
               forced_sel: Tensor | None = None):
    H, W = imgs.shape[-2:]
    depths = [disp_to_depth(resize_bilinear(d, (H, W)), min_depth, max_depth) for d in disps]
    l_rec, d_rec = image_recon(depths, imgs, supp_imgs, Ts, Ks, use_min, use_automask, noise, forced_sel=forced_sel)
    l_sm, d_sm = disp_smooth(disps, imgs, use_edges)
    loss = w_recon*l_rec + w_smooth*l_sm
    return loss, {'loss_img_recon': l_rec, 'loss_disp_smooth': l_sm, 'depth_up': depths, **d_rec, **d_sm}
