"""Photometric reconstruction loss — host-side mirror of `src/losses/{photometric,reconstruction}.py` (reference).

`ReconstructionLoss` keeps the reference's constructor and `forward` / `compute_photo` contracts. The training hot path
does not call `forward` on already-warped images: `slowtv_monodepth_b200.handlers.image_recon` hands the un-warped support
frames, depths and poses to the fused libstv kernel (`fused()` below), so the warped frames never exist in HBM; `forward`
(stv_recon_fwd / stv_recon_bwd) serves every other caller of the registered class.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as F_

__all__ = ['ReconstructionLoss', 'RegressionLoss']


class ReconstructionLoss(nn.Module):
    """Reference: src/losses/reconstruction.py:13-126.

    :param loss_name: 'ssim' (0.85 SSIM + 0.15 L1), 'l1' or 'l2' (the feature-space distance of `feat_recon`).
    :param use_min: minimum reprojection over the support frames instead of the mean.
    :param use_automask: mask pixels whose un-warped support frame already explains the target better.
    :param mask_name: None | 'explainability' | 'uncertainty' weighting of the per-frame errors (reconstruction.py:46-57).

    The single-pass fused kernel (`fused`) covers the KBR hot path: 3-channel frames, ssim | l1, no weighting mask. Everything
    else the class is registered for ('l2', masks, C-channel features) runs on the general kernels behind `forward`
    (stv_recon_ex_fwd / stv_recon_ex_bwd) after a stand-alone warp (`handlers.image_recon` picks the route).
    """
    def __init__(self, loss_name: str = 'ssim', use_min: bool = False, use_automask: bool = False, mask_name: str | None = None):
        super().__init__()
        if mask_name not in {'explainability', 'uncertainty', None}: raise ValueError(f'Invalid mask type: {mask_name}')
        if loss_name not in {'ssim', 'l1', 'l2'}: raise KeyError(loss_name)  # the reference indexes a dict (reconstruction.py:37-41)
        self.loss_name, self.use_min, self.use_automask, self.mask_name = loss_name, use_min, use_automask, mask_name
        self.noise_seed = 0x5107  # Base seed of the in-kernel tie-break noise (reconstruction.py:72 draws randn_like per call).
        # Device-side call counter added to the seed: the kernels advance it themselves, so eager calls AND replays of a captured
        # CUDA graph draw fresh noise every step. Created on first use on the inputs' device (before any graph capture: the
        # runners warm up eagerly first). Tests set it (`noise_step.fill_(k)`) to reproduce a particular draw.
        self.noise_step: Tensor | None = None

    @property
    def fusable(self) -> bool:
        """The single-pass warp + loss kernel serves this configuration (given 3-channel frames and no mask tensor)."""
        return self.loss_name in {'ssim', 'l1'} and self.mask_name is None

    def compute_photo(self, pred: Tensor, target: Tensor, mask: Tensor | None = None) -> Tensor:
        """pred (*n,b,c,h,w), target (b,c,h,w), mask (b,n,h,w) -> (b,1,h,w). Forward only (used for automasks and `depth_regr`)."""
        if self.mask_name and mask is None: raise ValueError("Must provide a 'mask' when masking...")
        if self.fusable and target.shape[1] == 3: return F_.photo_error(pred, target, loss_name=self.loss_name, use_min=self.use_min)
        return F_.photo_error_ex(pred, target, mask, loss_name=self.loss_name, use_min=self.use_min, mask_name=self.mask_name)

    def fused(self, depths: list[Tensor], target: Tensor, source: Tensor, T: Tensor, K: Tensor, K_inv: Tensor | None = None,
              noise: Tensor | None = None, want_warp: bool = False, from_disp: tuple | None = None):
        """Warp + loss in one kernel. -> (loss, {'automask': (S,b,1,H,W) bool}, sel, warp0).
        from_disp = (min_depth, max_depth): `depths` are the network's low-resolution sigmoid disparities; up-sampling and
        disparity -> depth happen inside the kernel."""
        if self.use_automask and noise is None: self._step_counter(target)
        loss, sel, warp0 = F_.photo_loss(depths, target, source, T, K, K_inv, loss_name=self.loss_name, use_min=self.use_min,
                                         use_automask=self.use_automask, noise=noise,
                                         noise_seed=self.noise_seed if self.use_automask else 0,
                                         noise_step=self.noise_step if (self.use_automask and noise is None) else None,
                                         want_warp=want_warp, **({} if from_disp is None else dict(
                                             disp_size=tuple(target.shape[-2:]), min_depth=from_disp[0], max_depth=from_disp[1])))
        ld = {'automask': (sel != 255).unsqueeze(2)} if self.use_automask else {}
        self.last_sel = sel  # (S,b,H,W) uint8 per-pixel decisions of the most recent call (diagnostics / parity tests)
        return loss, ld, sel, warp0

    def _step_counter(self, like: Tensor) -> Tensor:
        if self.noise_step is None or self.noise_step.device != like.device:
            self.noise_step = torch.zeros(1, dtype=torch.int64, device=like.device)
        return self.noise_step

    def forward(self, pred: Tensor, target: Tensor, source: Tensor | None = None, mask: Tensor | None = None, *,
                noise: Tensor | None = None):
        """Reference contract (reconstruction.py:98-126) on ALREADY WARPED frames: pred (*n,b,c,h,w), target (b,c,h,w),
        source (*n,b,c,h,w), mask (b,n,h,w) -> (loss, {'automask': (b,1,h,w) bool}); differentiable in `pred` and `mask`.
        The training hot path does not come through here — `handlers.image_recon` fuses the warp into the loss kernel — but
        everything else that calls the registered class directly does (the virtual-stereo branch, src/core/trainer.py:394-399;
        `feat_recon` / masked configurations via `handlers.image_recon`'s general route)."""
        if self.use_automask and source is None: raise ValueError("Must provide the original 'source' images when automasking...")
        draw = self.use_automask and noise is None
        if not (self.fusable and target.shape[1] == 3):
            loss, sel = F_.recon_loss_ex(pred, target, source if self.use_automask else None, mask, loss_name=self.loss_name,
                                         use_min=self.use_min, use_automask=self.use_automask, mask_name=self.mask_name, noise=noise,
                                         noise_seed=self.noise_seed if draw else 0, noise_step=self._step_counter(target) if draw else None)
            return loss, ({'automask': (sel < 128).unsqueeze(1)} if self.use_automask else {})
        loss, sel = F_.recon_loss(pred, target, source if self.use_automask else None, loss_name=self.loss_name, use_min=self.use_min,
                                  use_automask=self.use_automask, noise=noise, noise_seed=self.noise_seed if draw else 0,
                                  noise_step=self._step_counter(target) if draw else None)
        return loss, ({'automask': (sel != 255).unsqueeze(1)} if self.use_automask else {})


class RegressionLoss(nn.Module):
    """Reference: src/losses/regression.py:41-75 — the class registered as `depth_regr` and `stereo_const`.

    :param loss_name: 'l1' | 'log_l1' | 'berhu' (dynamic threshold 0.2 max|pred - target|).
    :param invert: convert both inputs with `to_inv` first (depths -> disparities).
    :param use_automask: read by `handlers.depth_regr` (the DepthHints automask is computed there, as in the reference).
    """
    def __init__(self, loss_name: str = 'berhu', invert: bool = False, use_automask: bool = False):
        super().__init__()
        if loss_name not in {'l1', 'log_l1', 'berhu'}: raise KeyError(loss_name)
        self.loss_name, self.invert, self.use_automask = loss_name, invert, use_automask

    def forward(self, pred: Tensor, target: Tensor, mask: Tensor | None = None):
        """-> (loss, {'err_regr': mask*err, 'mask_regr': mask}); differentiable in `pred` and `target` (stv_regr_fwd / stv_regr_bwd)."""
        if mask is None: mask = torch.ones_like(target)
        loss, err = F_.regr_loss(pred, target, mask, loss_name=self.loss_name, invert=self.invert)
        return loss, {'err_regr': err, 'mask_regr': mask}
