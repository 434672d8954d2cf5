"""Synthetic video-triplet batches in the reference's batch contract.

The reference datasets emit `(x, y, m)` (src/datasets/base_mde.py:158-176, documented at src/core/trainer.py:124-156):
  x['imgs']       (b,3,H,W)   ImageNet-standardised target frames (src/datasets/base_mde.py:276-281)
  x['supp_imgs']  (n,b,3,H,W) standardised support frames
  x['supp_idxs']  (n,)        int offsets of each support frame w.r.t. the target
  y['imgs'], y['supp_imgs']   the same frames, raw in [0,1]
  y['K']          (b,4,4)     pixel-unit intrinsics (KITTI-like: fx=.58W, fy=1.92H, cx=.5W, cy=.5H; kitti_raw.py:76-81)

There is no dataset access in the benchmark environment, so the generator draws a smooth random texture per sample and
shifts it by a few pixels per support frame: that makes min-reprojection and the automask both fire, as real video does.
Everything is generated on the CPU with a seeded `torch.Generator` so that every rank / device sees reproducible data.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

__all__ = ['supp_offsets', 'make_frames', 'make_batch', 'make_loss_inputs', 'kitti_like_K', 'standardize']


def kitti_like_K(b: int, shape: tuple[int, int], dtype=torch.float32) -> torch.Tensor:
    H, W = shape
    K = torch.eye(4, dtype=dtype).repeat(b, 1, 1)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = 0.58*W, 1.92*H, 0.5*W, 0.5*H
    return K


def supp_offsets(n: int) -> list[int]:
    """Support-frame offsets used for `n` support frames ([-1, 1] is the reference default, cfg/kbr/default.yaml:33)."""
    return {1: [1], 2: [-1, 1], 4: [-2, -1, 1, 2]}.get(n, list(range(1, n + 1)))


def standardize(x: torch.Tensor) -> torch.Tensor:
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype, device=x.device).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype, device=x.device).view(3, 1, 1)
    return (x - mean)/std


def make_frames(b: int, n: int, shape: tuple[int, int], seed: int = 0, dtype=torch.float32):
    """-> imgs (b,3,H,W), supp_imgs (n,b,3,H,W), both raw in [0,1]."""
    H, W = shape
    g = torch.Generator().manual_seed(seed)
    pad = 8
    coarse = torch.rand(b, 3, (H + 2*pad + 7)//8 + 1, (W + 2*pad + 7)//8 + 1, generator=g)
    base = F.interpolate(coarse, size=(H + 2*pad, W + 2*pad), mode='bilinear', align_corners=False)
    base = (base + 0.05*torch.rand(base.shape, generator=g)).clamp(0, 1)

    def crop(dy, dx): return base[..., pad + dy:pad + dy + H, pad + dx:pad + dx + W]

    imgs = crop(0, 0)
    supp = []
    for idx in supp_offsets(n):
        s = crop(max(-pad, min(pad, idx)), max(-pad, min(pad, 2*idx)))  # Frame `idx`: scene moved by (idx, 2 idx) px.
        s = (s + 0.02*(torch.rand(s.shape, generator=g) - 0.5)).clamp(0, 1)
        supp.append(s)
    return imgs.contiguous().to(dtype), torch.stack(supp).contiguous().to(dtype)


def make_batch(b: int, n: int, shape: tuple[int, int], seed: int = 0, device='cpu', pin: bool = False):
    """Full `(x, y, m)` batch in the reference's contract."""
    imgs, supp = make_frames(b, n, shape, seed)
    idxs = supp_offsets(n)
    x = {'imgs': standardize(imgs), 'supp_imgs': standardize(supp), 'supp_idxs': torch.tensor(idxs)}
    y = {'imgs': imgs, 'supp_imgs': supp, 'K': kitti_like_K(b, shape)}
    m = {}
    if pin:
        x = {k: v.pin_memory() for k, v in x.items()}
        y = {k: v.pin_memory() for k, v in y.items()}
    if str(device) != 'cpu':
        x = {k: (v.to(device, non_blocking=True) if k != 'supp_idxs' else v) for k, v in x.items()}
        y = {k: v.to(device, non_blocking=True) for k, v in y.items()}
    return x, y, m


def make_loss_inputs(b: int, n: int, S: int, shape: tuple[int, int], seed: int = 0, dtype=torch.float32,
                     learn_K: bool = False):
    """Inputs of the loss stack alone: multi-scale sigmoid disparities, frames, axis-angle/translation at the
    PoseNet output scale (src/networks/pose.py:128), intrinsics, and the automask tie-break noise.

    The pose is chosen so that, at the median depth, the reprojection roughly undoes the (idx, 2 idx) px scene shift of
    `make_frames`; with the disparity spread this leaves a healthy mix of pixels where (a) one or the other support frame
    wins the min-reprojection and (b) the automask fires or not."""
    H, W = shape
    imgs, supp = make_frames(b, n, shape, seed, dtype)
    g = torch.Generator().manual_seed(seed + 1)
    disps = []
    coarse = torch.rand(b, 1, max(H//32, 2), max(W//32, 2), generator=g, dtype=torch.float32)
    for s in range(S):
        h, w = H//2**s, W//2**s
        d = F.interpolate(coarse, size=(h, w), mode='bilinear', align_corners=False)
        d = (0.3 + 0.4*d + 0.05*torch.rand(b, 1, h, w, generator=g)).clamp(1e-3, 1 - 1e-3)
        disps.append(d.to(dtype))
    K = kitti_like_K(b, shape, torch.float32)
    if learn_K:
        K[:, 0, 0] *= 1 + 0.05*torch.randn(b, generator=g)
        K[:, 1, 1] *= 1 + 0.05*torch.randn(b, generator=g)
        K[:, 0, 2] += torch.randn(b, generator=g)
        K[:, 1, 2] += torch.randn(b, generator=g)
    d0 = 1/(9.99*0.5 + 0.01)
    idx = torch.tensor(supp_offsets(n), dtype=torch.float32).view(n, 1)
    t = 0.35*torch.stack([-2*idx*d0/K[:, 0, 0], -idx*d0/K[:, 1, 1], torch.zeros(n, b)], -1)  # (n, b, 3)
    t = (t + 0.0005*torch.randn(n, b, 3, generator=g)).to(dtype)
    aa = (0.002*torch.randn(n, b, 3, generator=g)).to(dtype)
    noise = torch.randn(S*b, 1, H, W, generator=g).to(dtype)
    return {'disps': disps, 'imgs': imgs, 'supp_imgs': supp, 'aa': aa, 't': t, 'K': K.to(dtype), 'noise': noise}
