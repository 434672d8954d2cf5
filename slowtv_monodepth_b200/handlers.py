"""Loss handlers — host-side mirror of `src/core/handlers.py` (reference) for the KBR loss set."""
from __future__ import annotations

import torch
from torch import Tensor

from .losses import ReconstructionLoss, RegressionLoss
from .regularizers import SmoothReg

__all__ = ['image_recon', 'feat_recon', 'stereo_const', 'depth_regr', 'disp_smooth', 'feat_smooth', 'disp_occ', 'disp_mask']


def _disparity_sources(depths: dict[int, Tensor], size) -> tuple | None:
    """(disparities, (min_depth, max_depth)) when every depth map is the tagged output of `functional.disp_to_depth` for this
    image size and one depth range — the loss kernel then starts from the disparities themselves."""
    tags = [getattr(d, '_stv_src', None) for d in depths.values()]
    if not tags or any(t is None for t in tags): return None
    if any(t[1] != tuple(size) or t[2:] != tags[0][2:] for t in tags): return None
    return [t[0] for t in tags], tags[0][2:]


def image_recon(crit: ReconstructionLoss, synth, depths: dict[int, Tensor], masks, imgs: Tensor, supp_imgs: Tensor,
                Ts: Tensor, Ks: Tensor, *, noise: Tensor | None = None, want_warp: bool = True,
                disps: dict[int, Tensor] | None = None, depth_range: tuple | None = None):
    """Reference: src/core/handlers.py:14-67.

    Same arguments and return contract. `synth` is accepted for signature compatibility and unused: the reference expands
    depths / frames / poses to n*S*b virtual images (189 MB of copies at b=8, 384x640) and warps them through ~60 ATen
    launches; here one kernel reads the originals once per (scale, support) and never materialises the warped frames.
    -> (loss, {'supp_imgs_warp': (n,b,3,H,W) at scale 0, 'automask': (b,1,H,W) bool at scale 0})

    Beyond the reference's arguments: `disps` + `depth_range=(min_depth, max_depth)` hand over the network's low-resolution
    disparities instead of `depths` (the kernel up-samples and converts them itself, trainer.py:320-321). When `depths` are the
    tagged outputs of `functional.disp_to_depth` (what the installed `forward_postprocess` produces) the same shortcut is taken
    automatically, so the reference's own `forward_loss` call site (trainer.py:389-393) gets it unchanged.
    """
    if masks is not None or not crit.fusable or imgs.shape[1] != 3:
        return _image_recon_general(crit, depths, masks, imgs, supp_imgs, Ts, Ks, noise=noise)
    if disps is None and crit.use_min and depths is not None:
        found = _disparity_sources(depths, imgs.shape[-2:])
        if found is not None: disps, depth_range = dict(zip(depths, found[0])), found[1]
    if disps is not None and (crit.use_min or depths is None):
        loss, ld, sel, warp0 = crit.fused(list(disps.values()), imgs, supp_imgs, Ts, Ks, noise=noise, want_warp=want_warp,
                                          from_disp=tuple(depth_range or (None, None)))
    else:
        keys = list(depths)
        loss, ld, sel, warp0 = crit.fused([depths[k] for k in keys], imgs, supp_imgs, Ts, Ks, noise=noise, want_warp=want_warp)
    out = {k: v[0] for k, v in ld.items()}  # Only scale 0 (handlers.py:64-65).
    if want_warp: out['supp_imgs_warp'] = warp0
    return loss, out


def _image_recon_general(crit: ReconstructionLoss, depths: dict[int, Tensor], masks, imgs: Tensor, supp_imgs: Tensor, Ts: Tensor,
                         Ks: Tensor, *, noise: Tensor | None = None):
    """The reference's own formulation (handlers.py:43-66) for what the single-pass kernel does not cover — predicted weighting
    masks, loss_name 'l2', C-channel feature maps: one stand-alone warp of the (n, S, b)-expanded batch (stv_view_synth, any
    channel count) followed by the general loss kernels (`ReconstructionLoss.forward`)."""
    from . import functional as F_
    n, S = supp_imgs.shape[0], len(depths)
    b, c, H, W = imgs.shape
    depth = torch.stack(list(depths.values()))                                            # (S, b, 1, H, W)
    depth = depth[None].expand(n, *depth.shape).reshape(n*S*b, 1, H, W)
    inp = supp_imgs[:, None].expand(n, S, b, c, H, W)
    T = Ts[:, None].expand(n, S, b, 4, 4).reshape(n*S*b, 4, 4)
    K = Ks[None, None].expand(n, S, b, 4, 4).reshape(n*S*b, 4, 4)
    warp = F_.view_synth(inp.reshape(n*S*b, c, H, W).contiguous(), depth.contiguous(), T.contiguous(), K.contiguous())[0]
    warp = warp.view(n, S*b, c, H, W)
    tgt = imgs[None].expand(S, b, c, H, W).reshape(S*b, c, H, W)
    src = inp.reshape(n, S*b, c, H, W)
    m = None
    if masks is not None: m = torch.stack(list(masks.values())).flatten(0, 1)            # (S*b, n, H, W)
    if noise is not None: noise = noise.reshape(S*b, 1, H, W)
    loss, ld = crit(warp, tgt.contiguous(), source=src.contiguous(), mask=m, noise=noise)
    out = {k: v.unflatten(0, (S, b))[0] for k, v in ld.items()}                          # Only scale 0 (handlers.py:63-64).
    out['supp_imgs_warp'] = warp.view(n, S, b, c, H, W)[:, 0]
    return loss, out


def feat_recon(crit: ReconstructionLoss, synth, depths: dict[int, Tensor], masks, feats, supp_feats, Ts: Tensor, Ks: Tensor):
    """Reference: src/core/handlers.py:70-119 — feature-space reconstruction at the highest-resolution depth map with the x4
    down-sampled encoder features, no gradient into the features. -> (loss, {'supp_feats_warp': (n,b,c,H,W)})."""
    from . import functional as F_
    if isinstance(feats, list): feats, supp_feats = feats[-4], supp_feats[-4]               # [*2, 4, 8, 16, 32] -> 4
    feats, supp_feats = feats.detach(), supp_feats.detach()
    size = tuple(depths[0].shape[-2:])
    with torch.no_grad():                                                                  # ops.interpolate_like(mode='bilinear')
        feats = F_.resample_bilinear(feats.contiguous(), size, mode='interp')
        supp_feats = F_.resample_bilinear(supp_feats.contiguous(), size, mode='interp')
    masks = {0: masks[0]} if masks is not None else None
    loss, ld = image_recon(crit, synth, {0: depths[0]}, masks, feats, supp_feats, Ts, Ks)
    return loss, {'supp_feats_warp': ld.pop('supp_imgs_warp')}


def stereo_const(crit: RegressionLoss, synth, disps: dict[int, Tensor], depths: dict[int, Tensor], disps_stereo: dict[int, Tensor],
                 depths_stereo: dict[int, Tensor], T_stereo: Tensor, K: Tensor):
    """Reference: src/core/handlers.py:151-198 — virtual stereo consistency: each view's disparities warped into the other view
    (stv_view_synth on the 1-channel maps) and regressed onto that view's own prediction (stv_regr_fwd/bwd; gradients reach both).
    -> (loss, {'disps_warp', 'stereo_disps_warp'} of the first scale)."""
    from . import functional as F_
    S = len(disps)
    st = lambda d: torch.stack(list(d.values())).flatten(0, 1)                             # (S*b, 1, H, W)
    d, z, ds, zs = st(disps), st(depths), st(disps_stereo), st(depths_stereo)
    b = d.shape[0]//S
    T = T_stereo[None].expand(S, b, 4, 4).reshape(S*b, 4, 4)
    Kx = K[None, None].expand(2, S, b, 4, 4).reshape(2*S*b, 4, 4)
    all_disps = torch.cat((ds, d))
    warp = F_.view_synth(all_disps.contiguous(), torch.cat((z, zs)).contiguous(), torch.cat((T, F_.inv4x4(T.contiguous()))).contiguous(),
                         Kx.contiguous())[0]
    loss, _ = crit(all_disps, warp)
    stereo_warp, disp_warp = warp.chunk(2)
    return loss, {'disps_warp': disp_warp.unflatten(0, (S, b))[0], 'stereo_disps_warp': stereo_warp.unflatten(0, (S, b))[0]}


def depth_regr(crit: RegressionLoss, synth, photo, depths: dict[int, Tensor], targets: Tensor, imgs: Tensor, supp_imgs: Tensor,
               Ts: Tensor, Ks: Tensor):
    """Reference: src/core/handlers.py:201-259 — proxy depth regression with the DepthHints automask (the hint must explain the
    target better than the prediction: photo(warp by depth) > photo(warp by hint)). `photo` = `ReconstructionLoss.compute_photo`.
    -> (loss, {'mask_regr'} of the first scale) — the reference rebinds its dict on the last line (handlers.py:257-258), so the
    `automask_hints` entry it builds never leaves the function; neither does it here."""
    from . import functional as F_
    S = len(depths)
    b, c, H, W = imgs.shape
    im = imgs[None].expand(S, b, c, H, W).reshape(S*b, c, H, W).contiguous()
    z = torch.stack(list(depths.values())).flatten(0, 1)                                  # (S*b, 1, H, W)
    tg = targets[None].expand(S, *targets.shape).reshape(S*b, 1, H, W).contiguous()
    masks = tg > 0
    if crit.use_automask:
        n = supp_imgs.shape[0]
        sup = supp_imgs[:, None].expand(n, S, b, c, H, W).reshape(n*S*b, c, H, W).contiguous()
        T = Ts[:, None].expand(n, S, b, 4, 4).reshape(n*S*b, 4, 4).contiguous()
        Kx = Ks[None, None].expand(n, S, b, 4, 4).reshape(n*S*b, 4, 4).contiguous()
        with torch.no_grad():   # the comparison is a boolean mask: no gradient leaves it
            rep = lambda x: x[None].expand(n, *x.shape).reshape(n*S*b, 1, H, W).contiguous()
            hints_warp = F_.view_synth(sup, rep(tg), T, Kx)[0].view(n, S*b, c, H, W)
            imgs_warp = F_.view_synth(sup, rep(z.detach()), T, Kx)[0].view(n, S*b, c, H, W)
            automask = photo(imgs_warp, im) > photo(hints_warp, im)
        masks = masks & automask
    loss, out = crit(z, tg, masks)
    return loss, {'mask_regr': out['mask_regr'].unflatten(0, (S, b))[0]}


def disp_smooth(crit: SmoothReg, disps: dict[int, Tensor], imgs: Tensor, *, want_maps: bool = True):
    """Reference: src/core/handlers.py:262-281. -> (loss, {'disp_grad', 'image_grad'} of the first scale)."""
    return crit.multi_scale(disps, imgs, want_maps=want_maps)


def feat_smooth(crit, feats, imgs: Tensor, supp_feats, supp_imgs: Tensor):
    """Reference: src/core/handlers.py:284-312 — FeatPeakReg / FeatSmoothReg over the multi-scale encoder features of the target and of
    the support frames, the images resized to each feature map (stv_resample_bilinear), loss_s / 2**s averaged. -> (loss, {})."""
    from . import functional as F_
    def one(fs, im):
        ls = [crit(f, F_.resample_bilinear(im, tuple(f.shape[-2:]), mode='interp') if f.shape[-2:] != im.shape[-2:] else im)[0]/2**s
              for s, f in enumerate(fs)]
        return sum(ls)/len(ls)
    loss = one(feats, imgs)
    loss = loss + one([f.flatten(0, 1) for f in supp_feats], supp_imgs.flatten(0, 1).contiguous())
    return loss, {}


def _per_scale_mean(crit, maps: dict[int, Tensor]):
    ls = {s: crit(m) for s, m in maps.items()}
    loss = sum(v[0] for v in ls.values())/len(ls)
    return loss, ls[0][1] if 0 in ls else next(iter(ls.values()))[1]   # loss dict of the first scale


def disp_occ(crit, disps: dict[int, Tensor]):
    """Reference: src/core/handlers.py:315-329 — OccReg per scale, averaged. -> (loss, {})."""
    return _per_scale_mean(crit, disps)


def disp_mask(crit, masks: dict[int, Tensor]):
    """Reference: src/core/handlers.py:332-346 — MaskReg per scale, averaged. -> (loss, {})."""
    return _per_scale_mean(crit, masks)
