"""Edge-aware disparity smoothness — host-side mirror of `src/regularizers/smooth.py` (reference)."""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor

from . import functional as F_

__all__ = ['SmoothReg', 'FeatPeakReg', 'FeatSmoothReg', 'MaskReg', 'OccReg']


class SmoothReg(nn.Module):
    """Reference: src/regularizers/smooth.py:51-97. The first-order variant (the KBR configuration) runs on the multi-scale hot-path
    kernels (stv_smooth_fwd/bwd); `use_laplacian` / `use_blur` run scale by scale on the general ones (stv_smooth_ex_fwd/bwd)."""
    def __init__(self, use_edges: bool = False, use_laplacian: bool = False, use_blur: bool = False):
        super().__init__()
        self.use_edges, self.use_laplacian, self.use_blur = use_edges, use_laplacian, use_blur

    @property
    def general(self) -> bool:
        return self.use_laplacian or self.use_blur

    def forward(self, disp: Tensor, img: Tensor):
        """disp (b,1,h,w), img (b,3,h,w) at the same resolution -> (loss, {'disp_grad', 'image_grad'})."""
        if disp.shape[-2:] != img.shape[-2:]: raise ValueError(f'Non-matching shapes. ({tuple(disp.shape)} vs. {tuple(img.shape)})')
        if self.general:
            loss, dg, ig = F_.smooth_loss_ex(disp, img, use_edges=self.use_edges, use_laplacian=self.use_laplacian, use_blur=self.use_blur)
            return loss, {'disp_grad': dg, 'image_grad': ig}
        loss, dg, ig = F_.smooth_loss([disp], img, scales=[0], use_edges=self.use_edges, want_maps=True)
        return loss, {'disp_grad': dg, 'image_grad': ig}

    def multi_scale(self, disps: dict[int, Tensor], imgs: Tensor, want_maps: bool = True):
        """All scales in one launch set: mean_s(loss_s / 2**s) with the image resized in-kernel (handlers.py:278-279)."""
        keys = list(disps)
        if self.general:   # handlers.py:276-280 verbatim: per scale, image resized to the disparity, loss_s / 2**s averaged
            ls, maps = [], {}
            for k in keys:
                img_k = imgs if disps[k].shape[-2:] == imgs.shape[-2:] else F_.resample_bilinear(imgs, tuple(disps[k].shape[-2:]), mode='interp')
                l, maps[k] = self.forward(disps[k], img_k)
                ls.append(l/2**k)
            return sum(ls)/len(ls), (maps.get(0, maps[keys[0]]) if want_maps else {})   # loss dict of the first scale (handlers.py:280)
        loss, dg, ig = F_.smooth_loss([disps[k] for k in keys], imgs, scales=keys, use_edges=self.use_edges, want_maps=want_maps)
        return loss, ({'disp_grad': dg, 'image_grad': ig} if want_maps else {})


class FeatPeakReg(nn.Module):
    """Reference: src/regularizers/smooth.py:100-136 (`feat_peaky`): encourage first-order feature gradients (negative mean)."""
    def __init__(self, use_edges: bool = False):
        super().__init__()
        self.use_edges = use_edges

    def forward(self, feat: Tensor, img: Tensor):
        loss, fg = F_.feat_reg(feat, img, order=1, use_edges=self.use_edges)
        return loss, {'feat_grad': fg}


class FeatSmoothReg(nn.Module):
    """Reference: src/regularizers/smooth.py:139-176 (`feat_smooth`): penalise second-order feature gradients (xx, yy, xy, yx)."""
    def __init__(self, use_edges: bool = False):
        super().__init__()
        self.edge_aware = use_edges

    def forward(self, feat: Tensor, img: Tensor):
        loss, fg = F_.feat_reg(feat, img, order=2, use_edges=self.edge_aware)
        return loss, {'feat_grad': fg}


class MaskReg(nn.Module):
    """Reference: src/regularizers/mask.py:11-30 (`disp_mask`): binary cross-entropy of the explainability mask against 1."""
    def forward(self, x: Tensor):
        return F_.bce_to_one(x), {}


class OccReg(nn.Module):
    """Reference: src/regularizers/occlusion.py:9-40 (`disp_occ`): +-mean disparity (`invert` encourages foreground instead)."""
    def __init__(self, invert: bool = False):
        super().__init__()
        self.invert, self._sign = invert, (-1 if invert else 1)

    def forward(self, x: Tensor):
        return F_.mean_reg(x, self._sign), {}
