"""Edge-aware disparity smoothness — host-side mirror of `src/regularizers/smooth.py` (reference)."""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor

from . import functional as F_

__all__ = ['SmoothReg']


class SmoothReg(nn.Module):
    """Reference: src/regularizers/smooth.py:51-97 (first-order variant; Laplacian / blur are ablation-only, rejected)."""
    def __init__(self, use_edges: bool = False, use_laplacian: bool = False, use_blur: bool = False):
        super().__init__()
        if use_laplacian or use_blur:
            raise NotImplementedError('use_laplacian / use_blur are outside the B200 hot path (SURVEY 8a row 16).')
        self.use_edges, self.use_laplacian, self.use_blur = use_edges, use_laplacian, use_blur

    def forward(self, disp: Tensor, img: Tensor):
        """disp (b,1,h,w), img (b,3,h,w) at the same resolution -> (loss, {'disp_grad', 'image_grad'})."""
        if disp.shape[-2:] != img.shape[-2:]: raise ValueError(f'Non-matching shapes. ({tuple(disp.shape)} vs. {tuple(img.shape)})')
        loss, dg, ig = F_.smooth_loss([disp], img, scales=[0], use_edges=self.use_edges, want_maps=True)
        return loss, {'disp_grad': dg, 'image_grad': ig}

    def multi_scale(self, disps: dict[int, Tensor], imgs: Tensor, want_maps: bool = True):
        """All scales in one launch set: mean_s(loss_s / 2**s) with the image resized in-kernel (handlers.py:278-279)."""
        keys = list(disps)
        loss, dg, ig = F_.smooth_loss([disps[k] for k in keys], imgs, scales=keys, use_edges=self.use_edges, want_maps=want_maps)
        return loss, ({'disp_grad': dg, 'image_grad': ig} if want_maps else {})
