"""Depth network — host-side mirror of `src/networks/depth.py` (reference) for the KBR configuration."""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor

from .decoder import MonodepthDecoder
from .encoders import create_encoder

__all__ = ['DepthNet']

DECODERS = {'monodepth': MonodepthDecoder}


class DepthNet(nn.Module):
    """Reference: src/networks/depth.py:17-156. Same constructor signature; `mask_name`, `use_virtual_stereo` and
    `use_stereo_blend` (not used by the KBR configs, cfg/kbr/default.yaml:3-11) are rejected loudly."""
    def __init__(self, enc_name: str = 'resnet18', pretrained: bool = True, dec_name: str = 'monodepth', out_scales=(0, 1, 2, 3),
                 mask_name=None, num_ch_mask=None, use_virtual_stereo: bool = False, use_stereo_blend: bool = False):
        super().__init__()
        if dec_name not in DECODERS: raise KeyError(f'Invalid decoder. ({dec_name} vs. {list(DECODERS)}')
        if mask_name is not None: raise KeyError(f'Invalid mask. ({mask_name}): mask prediction is outside the B200 hot path.')
        if use_virtual_stereo or use_stereo_blend: raise NotImplementedError('Virtual stereo is outside the B200 hot path.')
        self.enc_name, self.pretrained, self.dec_name = enc_name, pretrained, dec_name
        self.out_scales = [out_scales] if isinstance(out_scales, int) else list(out_scales)
        self.mask_name, self.num_ch_mask = mask_name, num_ch_mask
        self.use_virtual_stereo, self.use_stereo_blend = use_virtual_stereo, use_stereo_blend

        self.encoder = create_encoder(enc_name, in_chans=3, pretrained=pretrained)
        self.num_ch_enc, self.enc_sc = self.encoder.feature_info.channels(), self.encoder.feature_info.reduction()
        self.decoders = nn.ModuleDict({'disp': DECODERS[dec_name](
            num_ch_enc=self.num_ch_enc, enc_sc=self.enc_sc, upsample_mode='nearest', use_skip=True,
            out_sc=self.out_scales, out_ch=1, out_act='sigmoid')})

    def forward(self, x: Tensor) -> dict:
        feat = self.encoder(x)
        disp = self.decoders['disp'](feat)
        return {'depth_feats': feat, 'disp': dict(sorted(disp.items()))}
