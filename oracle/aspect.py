"""CPU restatement of the reference's aspect-ratio augmentation (src/core/aspect_ratio.py).  TEST INFRASTRUCTURE ONLY: imported by
tests/, oracle/make_golden_aspect.py and nothing else; the product path is slowtv_monodepth_b200/aspect_ratio.py (CUDA).

Pinning:
  * `sample_crop` (ref :100-126), `sample_resize` (:170-186), `crop_aug` (:69-97), `resize_aug` (:129-167), `aspect_ratio_aug`
    (:36-66), `centre_crop_K` / `resize_K` (src/tools/geometry.py:233-263) are checked against the reference's OWN functions run
    in this container under identical RNG seeds (tests/test_aspect_cpu.py, golden file tests/golden/aspect_cases.json made by
    oracle/make_golden_aspect.py).
  * `center_crop` is third-party arithmetic that is ABSENT here: kornia==0.6.10 (docker/environment.yml:118), call site
    aspect_ratio.py:82 `KT.center_crop(size, mode='bilinear', align_corners=False)`. It is restated from kornia 0.6.10's published
    algorithm (geometry/transform/crop2d.py center_crop -> crop_by_boxes -> crop_by_transform_mat -> imgwarp.warp_affine):
    the crop is a warp by the translation that maps the box [start, start+size-1] onto [0, size-1], evaluated through
    `normalize_homography` (pixel -> [-1,1] with the (size-1) convention), `F.affine_grid(align_corners=False)` and
    `F.grid_sample(bilinear, zeros, align_corners=False)`. Mixing the two conventions makes it a genuine resampling, not a
    copy:   ix(j) = (start + (j + 0.5) * (wd - 1)/wd) * W/(W - 1) - 0.5.   PARITY UNPINNED for this one function (no kornia
    here, no reference test); kornia additionally solves the 4-point homography in float32, whose rounding is not modelled.
"""
from __future__ import annotations

import random

import torch
import torch.nn.functional as F

LABELS = ['6/13', '9/16', '3/5', '2/3', '4/5', '1/1', '5/4', '4/3', '3/2', '14/9', '5/3', '16/9', '2/1', '24/10', '33/10', '18/5']
RATIOS = [int(a)/int(b) for a, b in (s.split('/') for s in LABELS)]


def sample_crop(shape, lo=0.5, hi=1.0):
    """ref :100-126 — ten candidate heights and widths, one ratio; candidates are (h, r*h) and (w/r, w); a valid one is drawn."""
    n = 10
    hs = torch.randint(int(shape[0]*lo), int(shape[0]*hi), (n,))
    ws = torch.randint(int(shape[1]*lo), int(shape[1]*hi), (n,))
    r = random.choice(RATIOS)
    hs, ws = torch.cat((hs, (ws/r).long())), torch.cat(((r*hs).long(), ws))
    valid = (hs >= 0) & (hs <= shape[0]) & (ws >= 0) & (ws <= shape[1])
    i = random.choice(valid.nonzero().squeeze())
    return (hs[i].item(), ws[i].item()), r


def sample_resize(shape, ref_shape, eps=0.8):
    """ref :170-186 — multiples of 32 with (at most) eps * the reference pixel count."""
    n, n_ref = shape[0]*shape[1], ref_shape[0]*ref_shape[1]
    r = (n_ref/n)**0.5
    res = [round(r*i/32)*32 for i in shape]
    while res[0]*res[1] > n_ref*eps: res = [i - 32 for i in res]
    return res


def crop_start(src: int, dst: int) -> int:
    """kornia 0.6.10 center_crop: start = int(src/2 - dst/2)."""
    return int(src/2 - dst/2)


def center_crop(x: torch.Tensor, size) -> torch.Tensor:
    """kornia 0.6.10 `center_crop(x, size, mode='bilinear', padding_mode='zeros', align_corners=False)` restated (see header)."""
    B, _, H, W = x.shape
    dh, dw = size
    sy, sx = crop_start(H, dh), crop_start(W, dw)
    dt = x.dtype
    M = torch.tensor([[1., 0., -sx], [0., 1., -sy], [0., 0., 1.]], dtype=dt)       # dst_pix <- src_pix

    def norm(h, w):  # kornia normal_transform_pixel
        return torch.tensor([[2./(w - 1), 0., -1.], [0., 2./(h - 1), -1.], [0., 0., 1.]], dtype=dt)

    dst_norm_trans_src_norm = norm(dh, dw) @ M @ torch.linalg.inv(norm(H, W))
    theta = torch.linalg.inv(dst_norm_trans_src_norm)[None, :2, :].expand(B, -1, -1)
    grid = F.affine_grid(theta, [B, x.shape[1], dh, dw], align_corners=False)
    return F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=False)


def centre_crop_K(K, new_shape, shape):
    K = K.clone()
    K[..., 0, 2] *= new_shape[1]/shape[1]
    K[..., 1, 2] *= new_shape[0]/shape[0]
    return K


def resize_K(K, new_shape, shape):
    K = K.clone()
    K[..., 0, :] *= new_shape[1]/shape[1]
    K[..., 1, :] *= new_shape[0]/shape[0]
    return K


def _apply(batch, fn):
    x, y, m = batch
    for d in (x, y):
        d['imgs'] = fn(d['imgs'])
        n, b = d['supp_imgs'].shape[:2]
        d['supp_imgs'] = fn(d['supp_imgs'].flatten(0, 1)).unflatten(0, (n, b))
    if 'depth' in y: y['depth'] = fn(y['depth'])
    return x, y, m


def crop_aug(batch, lo=0.5, hi=1.0):
    x, y, m = batch
    shape = tuple(x['imgs'].shape[-2:])
    crop_shape, ratio = sample_crop(shape, lo, hi)
    m.setdefault('augs', []).append(f'{list(shape)} -> {crop_shape} -> {LABELS[RATIOS.index(ratio)]}')
    _apply(batch, lambda t: center_crop(t, crop_shape))
    if 'K' in y: y['K'] = centre_crop_K(y['K'], crop_shape, shape)
    return batch


def resize_aug(batch, ref_shape, eps=0.8):
    x, y, m = batch
    new_shape = tuple(x['imgs'].shape[-2:])
    res_shape = sample_resize(new_shape, ref_shape, eps)
    m.setdefault('augs', []).append(str(res_shape))
    _apply(batch, lambda t: F.interpolate(t, size=res_shape, mode='bilinear', align_corners=False))
    if 'K' in y: y['K'] = resize_K(y['K'], res_shape, new_shape)
    return batch


def aspect_ratio_aug(batch, p=1.0, crop_min=0.5, crop_max=1.0, ref_shape=None):
    sh = tuple(batch[0]['imgs'].shape[-2:])
    if random.random() > p:
        return resize_aug(batch, ref_shape, eps=1) if ref_shape and tuple(ref_shape) != sh else batch
    ref_shape = ref_shape or sh
    batch = crop_aug(batch, crop_min, crop_max)
    return resize_aug(batch, ref_shape, eps=0.8)
