"""Depth / pose / intrinsics helpers and view synthesis — host-side mirror of `src/tools/geometry.py` (reference).

The heavy lifting (`ViewSynth`, disparity post-processing) runs in libstv kernels; the tiny per-sample matrix algebra
(`T_from_AAt`, `resize_K`, `K.inverse()`) stays in PyTorch autograd, as SURVEY 8a row 5 prescribes.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as F_

__all__ = ['eps', 'to_scaled', 'to_inv', 'T_from_AAt', 'resize_K', 'ViewSynth', 'upsample_to_depth']


def eps(x: Tensor | None = None) -> float:
    """Machine epsilon of the tensor's dtype (reference: src/tools/ops.py:63-66)."""
    return torch.finfo(torch.float32 if x is None else x.dtype).eps


def to_inv(depth: Tensor) -> Tensor:
    """Linear depth <-> disparity (reference: src/tools/geometry.py:86-90). Elementwise; differentiable."""
    return (depth > 0)/depth.clamp(min=eps(depth))


def to_scaled(disp: Tensor, min: float = 0.01, max: float | None = 100) -> tuple[Tensor, Tensor]:
    """Sigmoid disparity -> (scaled disparity, depth) (reference: src/tools/geometry.py:62-76)."""
    if min <= 0: raise ValueError(f'Min depth must be greater than 0. ({min})')
    if max and (max < min): raise ValueError(f'Max depth must be greater than min. ({max} vs. {min})')
    i_max, i_min = 1/min, (1/max) if max else 0
    disp = (i_max - i_min)*disp + i_min
    return disp, to_inv(disp)


def upsample_to_depth(disp: Tensor, size: tuple[int, int], min_depth: float | None, max_depth: float | None):
    """Fused `ops.interpolate_like(disp, imgs, 'bilinear')` + `to_depth` (reference: src/core/trainer.py:320-321).
    -> (disp_up, depth_up), one kernel each way."""
    return F_.disp_to_depth(disp, size, min_depth, max_depth)


def T_from_AAt(aa: Tensor, t: Tensor) -> Tensor:
    """Axis-angle + translation -> (*, 4, 4) transform (reference: src/tools/geometry.py:181-209; Rodrigues formula with
    axis = aa/max(|aa|, eps), src/tools/geometry.py:136-140)."""
    s1, s2 = aa.shape, t.shape
    if s1[-1] != 3: raise ValueError(f'Incorrect `axisangle` shape. ({s1} vs. (*, 3)')
    if s2[-1] != 3: raise ValueError(f'Incorrect `t` shape. ({s2} vs. (*, 3)')
    if s1 != s2: raise ValueError(f'Non-matching shapes. ({s1} vs. {s2}')

    angle = aa.norm(p=2, dim=-1, keepdim=True)
    x, y, z = (aa/angle.clip(min=eps(angle))).unbind(-1)
    o = torch.zeros_like(x)
    Wm = torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=-1).unflatten(-1, (3, 3))
    a = angle.unsqueeze(-1)
    R = torch.eye(3, dtype=aa.dtype, device=aa.device) + Wm*a.sin() + (Wm @ Wm)*(1 - a.cos())
    top = torch.cat([R, t.unsqueeze(-1)], dim=-1)
    bot = torch.zeros_like(top[..., :1, :])
    bot[..., 0, 3] = 1
    return torch.cat([top, bot], dim=-2)


def resize_K(K: Tensor, new_shape: tuple[int, int], shape: tuple[int, int] | None = None) -> Tensor:
    """Scale the intrinsics' first two rows with the image size (reference: src/tools/geometry.py:249-263)."""
    if shape is None: shape = (1, 1)
    sc = torch.ones(4, 1, dtype=K.dtype, device=K.device)
    sc[0], sc[1] = new_shape[1]/shape[1], new_shape[0]/shape[0]
    return K*sc


class ViewSynth(nn.Module):
    """Warp an image according to depth and pose (reference: src/tools/geometry.py:353-391).

    Same constructor and `forward` contract; the back-project / transform / project / grid-sample chain is a single
    libstv kernel each way, and no pixel-grid buffers are built (the reference rebuilds them on the CPU every step,
    src/core/trainer.py:168).
    """
    def __init__(self, shape: tuple[int, int]):
        super().__init__()
        self.shape = tuple(shape)

    def forward(self, input: Tensor, depth: Tensor, T: Tensor, K: Tensor, K_inv: Tensor | None = None):
        if tuple(input.shape[-2:]) != self.shape:
            raise ValueError(f'Input does not match the ViewSynth shape. ({tuple(input.shape[-2:])} vs. {self.shape})')
        return F_.view_synth(input, depth, T, K, K_inv)
