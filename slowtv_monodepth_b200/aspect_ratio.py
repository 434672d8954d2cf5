"""Aspect-ratio augmentation on the GPU — host-side mirror of src/core/aspect_ratio.py (reference), SURVEY 8(f) rank 1.

Same entry point and semantics as the reference's `aspect_ratio_aug(batch, p, crop_min, crop_max, ref_shape)` (bound with
`functools.partial` in `MonoDepthModule.__init__`, src/core/trainer.py:54-60, and applied to every training batch at
trainer.py:106): with probability `p` a centre crop of a randomly drawn aspect ratio followed by a resize to (a multiple of
32 with) at most 0.8x the reference pixel count; otherwise at most a plain resize to `ref_shape`. The batch is modified in
place; intrinsics follow the images (`centre_crop_K`, `resize_K`, src/tools/geometry.py:233-263).

What runs where:
  * the random draws consume Python's `random` and torch's CPU generator in the reference's order (aspect_ratio.py:53,
    113-114, 116, 120), so a seeded run picks the same crops and sizes as the reference;
  * the two image operations are ONE libstv kernel each (stv_resample_bilinear) per tensor, straight on the (…,3,H,W) buffers —
    no `torch.cat` of the four image tensors and `split` back (the reference's "speed up" copies every image twice more).
    `kornia.center_crop(bilinear, align_corners=False)` is a genuine resampling (kornia mixes the (size-1) pixel normalisation
    with align_corners=False sampling), reproduced through the kernel's affine mode; see oracle/aspect.py for the derivation.
There is no CPU path: host tensors raise.
"""
from __future__ import annotations

import random

import torch
from torch import Tensor

from . import functional as F_
from .geometry import resize_K

__all__ = ['aspect_ratio_aug', 'crop_aug', 'resize_aug', 'sample_crop', 'sample_resize', 'centre_crop_K', 'center_crop', 'LABELS', 'RATIOS']

LABELS = [
    '6/13', '9/16', '3/5', '2/3', '4/5', '1/1',                                            # portrait
    '5/4', '4/3', '3/2', '14/9', '5/3', '16/9', '2/1', '24/10', '33/10', '18/5',           # landscape
]
RATIOS = [int(s.split('/')[0])/int(s.split('/')[1]) for s in LABELS]
MULTIPLE = 32   # network stride: output sizes are multiples of it (aspect_ratio.py:181)
N_CANDIDATES = 10


def sample_crop(shape: tuple[int, int], min: float = 0.5, max: float = 1.0) -> tuple[tuple[int, int], float]:
    """Draw a centre-crop size with a random aspect ratio (aspect_ratio.py:100-126): candidate heights and widths in
    [min, max) of the image, each completed with the drawn ratio; one candidate that fits the image is chosen."""
    if max < min: raise ValueError(f'sample_crop: max < min ({max} vs. {min})')
    H, W = shape
    cand_h = torch.randint(int(H*min), int(H*max), (N_CANDIDATES,))
    cand_w = torch.randint(int(W*min), int(W*max), (N_CANDIDATES,))
    ratio = random.choice(RATIOS)
    hs = torch.cat((cand_h, (cand_w/ratio).long()))
    ws = torch.cat(((ratio*cand_h).long(), cand_w))
    fits = (hs >= 0) & (hs <= H) & (ws >= 0) & (ws <= W)
    pick = random.choice(fits.nonzero().squeeze())
    return (hs[pick].item(), ws[pick].item()), ratio


def sample_resize(shape: tuple[int, int], ref_shape: tuple[int, int], eps: float = 0.8) -> list[int]:
    """Size with the aspect ratio of `shape`, sides multiples of 32, at most eps x the pixels of `ref_shape` (aspect_ratio.py:170-186)."""
    budget = ref_shape[0]*ref_shape[1]
    scale = (budget/(shape[0]*shape[1]))**0.5
    out = [round(scale*side/MULTIPLE)*MULTIPLE for side in shape]
    while out[0]*out[1] > budget*eps: out = [side - MULTIPLE for side in out]
    return out


def centre_crop_K(K: Tensor, new_shape: tuple[int, int], shape: tuple[int, int] | None = None) -> Tensor:
    """Principal point follows a centre crop (src/tools/geometry.py:233-246)."""
    if shape is None: shape = (1, 1)
    sc = torch.ones(4, 4, dtype=K.dtype, device=K.device)
    sc[0, 2], sc[1, 2] = new_shape[1]/shape[1], new_shape[0]/shape[0]
    return K*sc


def _crop_axis(src: int, dst: int) -> tuple[float, float]:
    """Sample position of output index j along one axis of kornia 0.6.10's centre crop: a*j + b (pixel units)."""
    start = int(src/2 - dst/2)
    shrink, stretch = (dst - 1)/dst, src/(src - 1)
    return shrink*stretch, (start + 0.5*shrink)*stretch - 0.5


def center_crop(x: Tensor, size: tuple[int, int]) -> Tensor:
    """`kornia.geometry.transform.center_crop(x, size, mode='bilinear', align_corners=False)` on (..., H, W) device tensors."""
    H, W = x.shape[-2:]
    if H < 2 or W < 2 or size[0] < 1 or size[1] < 1: raise ValueError(f'center_crop: degenerate shapes {tuple(x.shape)} -> {size}')
    (ay, by), (ax, bx) = _crop_axis(H, size[0]), _crop_axis(W, size[1])
    return F_.resample_bilinear(x, size, mode='grid', ax=ax, bx=bx, ay=ay, by=by)


def _images(batch, fn) -> None:
    x, y, _ = batch
    for d in (x, y):
        d['imgs'] = fn(d['imgs'])
        d['supp_imgs'] = fn(d['supp_imgs'])   # (n,b,3,H,W): the kernel treats every leading axis as planes
    return


def crop_aug(batch, min: float = 0.5, max: float = 1.0):
    """Centre crop with a random aspect ratio (aspect_ratio.py:69-97)."""
    x, y, m = batch
    shape = tuple(x['imgs'].shape[-2:])
    crop_shape, ratio = sample_crop(shape, min, max)
    m.setdefault('augs', []).append(f'{list(shape)} -> {crop_shape} -> {LABELS[RATIOS.index(ratio)]}')
    fn = lambda t: center_crop(t, crop_shape)
    _images(batch, fn)
    if 'depth' in y: y['depth'] = fn(y['depth'])
    if 'depth_hints' in y: y['depth_hints'] = fn(y['depth_hints'])
    if 'K' in y: y['K'] = centre_crop_K(y['K'], crop_shape, shape)
    return x, y, m


def resize_aug(batch, ref_shape: tuple[int, int], eps: float = 0.8):
    """Resize to (at most) eps x the pixel count of `ref_shape` (aspect_ratio.py:129-167)."""
    x, y, m = batch
    new_shape = tuple(x['imgs'].shape[-2:])
    res_shape = sample_resize(new_shape, ref_shape, eps=eps)
    m.setdefault('augs', []).append(str(res_shape))
    fn = lambda t: F_.resample_bilinear(t, res_shape, mode='interp')
    _images(batch, fn)
    if 'depth' in y: y['depth'] = fn(y['depth'])
    if 'depth_hints' in y:
        raise RuntimeError('Geometric augmentation should not be combined with depth hints... Interpolating depth is not well defined.')
    if 'K' in y: y['K'] = resize_K(y['K'], res_shape, shape=new_shape)
    return x, y, m


@torch.no_grad()
def aspect_ratio_aug(batch, p: float = 1.0, crop_min: float = 0.5, crop_max: float = 1.0, ref_shape: tuple[int, int] | None = None):
    """Reference signature (aspect_ratio.py:36-66). In-place on `batch`; returns it."""
    sh = tuple(batch[0]['imgs'].shape[-2:])
    if random.random() > p:
        return resize_aug(batch, ref_shape, eps=1) if ref_shape and tuple(ref_shape) != sh else batch
    ref_shape = ref_shape or sh
    batch = crop_aug(batch, min=crop_min, max=crop_max)
    return resize_aug(batch, ref_shape=ref_shape, eps=0.8)
