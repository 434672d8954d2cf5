"""Loss handlers — host-side mirror of `src/core/handlers.py` (reference) for the KBR loss set."""
from __future__ import annotations

import torch
from torch import Tensor

from .losses import ReconstructionLoss
from .regularizers import SmoothReg

__all__ = ['image_recon', 'disp_smooth']


def image_recon(crit: ReconstructionLoss, synth, depths: dict[int, Tensor], masks, imgs: Tensor, supp_imgs: Tensor,
                Ts: Tensor, Ks: Tensor, *, noise: Tensor | None = None, want_warp: bool = True):
    """Reference: src/core/handlers.py:14-67.

    Same arguments and return contract. `synth` is accepted for signature compatibility and unused: the reference expands
    depths / frames / poses to n*S*b virtual images (189 MB of copies at b=8, 384x640) and warps them through ~60 ATen
    launches; here one kernel reads the originals once per (scale, support) and never materialises the warped frames.
    -> (loss, {'supp_imgs_warp': (n,b,3,H,W) at scale 0, 'automask': (b,1,H,W) bool at scale 0})
    """
    if masks is not None: raise ValueError('Predicted photometric masks are not supported by the B200 loss kernels.')
    keys = list(depths)
    loss, ld, sel, warp0 = crit.fused([depths[k] for k in keys], imgs, supp_imgs, Ts, Ks, noise=noise, want_warp=want_warp)
    out = {k: v[0] for k, v in ld.items()}  # Only scale 0 (handlers.py:64-65).
    if want_warp: out['supp_imgs_warp'] = warp0
    return loss, out


def disp_smooth(crit: SmoothReg, disps: dict[int, Tensor], imgs: Tensor, *, want_maps: bool = True):
    """Reference: src/core/handlers.py:262-281. -> (loss, {'disp_grad', 'image_grad'} of the first scale)."""
    return crit.multi_scale(disps, imgs, want_maps=want_maps)
