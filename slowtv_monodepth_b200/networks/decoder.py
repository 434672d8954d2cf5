"""Monodepth2-style disparity decoder — host-side mirror of `src/networks/decoders/monodepth.py` (reference).

Parameter layout matches the reference's `state_dict`: the reference keeps its layers in `self.decoder = nn.ModuleList(...)`
in creation order (monodepth.py:51-69), i.e. `decoder.{2*(4-i)+j}.conv.{weight,bias}` for `upconv_i_j` and
`decoder.{10+idx}.{weight,bias}` for `outconv_i`.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .. import _lib as L_
from .. import functional as F_

__all__ = ['MonodepthDecoder']

_ACT = {'sigmoid': torch.sigmoid, 'relu': F.relu, 'none': lambda x: x, None: lambda x: x}


class _ConvBlock(nn.Module):
    """3x3 reflect-padded convolution + ELU (reference `conv_block`, src/networks/decoders/utils.py:44-54)."""
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1, padding_mode='reflect')

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda: return F_.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous(), self.conv.weight, self.conv.bias, pad=1, reflect=True,
                                            act='elu').permute(0, 3, 1, 2)
        return L_.host_path(self, x)


class MonodepthDecoder(nn.Module):
    """Same constructor arguments and outputs as the reference class (monodepth.py:15-89)."""
    def __init__(self, num_ch_enc, enc_sc, upsample_mode: str = 'nearest', use_skip: bool = True, out_sc=(0, 1, 2, 3),
                 out_ch: int = 1, out_act: str = 'sigmoid'):
        super().__init__()
        if out_act not in _ACT: raise KeyError(f'Invalid activation key. ({out_act} vs. {tuple(_ACT.keys())}')
        self.num_ch_enc, self.enc_sc = list(num_ch_enc), list(enc_sc)
        self.upsample_mode, self.use_skip, self.out_sc, self.out_ch, self.out_act = upsample_mode, use_skip, list(out_sc), out_ch, out_act
        self.num_ch_dec = [16, 32, 64, 128, 256]

        layers, self._idx = [], {}
        for i in range(4, -1, -1):
            cin = self.num_ch_enc[-1] if i == 4 else self.num_ch_dec[i + 1]
            self._idx[f'upconv_{i}_0'] = len(layers); layers.append(_ConvBlock(cin, self.num_ch_dec[i]))
            cin = self.num_ch_dec[i]
            if self.use_skip and 2**i in self.enc_sc: cin += self.num_ch_enc[self.enc_sc.index(2**i)]
            self._idx[f'upconv_{i}_1'] = len(layers); layers.append(_ConvBlock(cin, self.num_ch_dec[i]))
        for i in self.out_sc:
            self._idx[f'outconv_{i}'] = len(layers)
            layers.append(nn.Conv2d(self.num_ch_dec[i], self.out_ch, 3, padding=1, padding_mode='reflect'))
        self.decoder = nn.ModuleList(layers)

    def layer(self, name: str) -> nn.Module:
        return self.decoder[self._idx[name]]

    def forward(self, feat: list[Tensor]) -> dict[int, Tensor]:
        if feat[-1].is_cuda and self.upsample_mode == 'nearest': return self.forward_nhwc(feat)
        return L_.host_path(self, feat, what='MonodepthDecoder' if not feat[-1].is_cuda else f'MonodepthDecoder(upsample_mode={self.upsample_mode!r})')

    def forward_nhwc(self, feat: list[Tensor]) -> dict[int, Tensor]:
        """The same network on channels-last tensors with libstv tcgen05 implicit-GEMM convolutions: reflection padding,
        nearest x2 upsampling and the skip concatenation are resolved inside the operand gather (no padded / upsampled /
        concatenated tensor is ever written), bias + ELU / sigmoid in the epilogue."""
        f = [t.permute(0, 2, 3, 1) for t in feat]
        out, act = {}, (None if self.out_act in ('none', None) else self.out_act)
        x = f[-1]
        for i in range(4, -1, -1):
            c0, c1 = self.layer(f'upconv_{i}_0').conv, self.layer(f'upconv_{i}_1').conv
            x = F_.conv2d_nhwc(x, c0.weight, c0.bias, pad=1, reflect=True, act='elu')
            if i == 4: self.first_node = x.grad_fn   # the decoder's last node in backward: its gradients are final after it
            skip = f[self.enc_sc.index(2**i)] if self.use_skip and 2**i in self.enc_sc else None
            x = F_.conv2d_nhwc(x, c1.weight, c1.bias, src2=skip, up1=True, pad=1, reflect=True, act='elu')
            if i in self.out_sc:
                h = self.layer(f'outconv_{i}')
                if F_.head3x3_supported(x.shape[-1], self.out_ch): o = F_.head3x3(x, h.weight, h.bias, act)  # per-pixel dot product
                else: o = F_.conv2d_nhwc(x, h.weight, h.bias, pad=1, reflect=True, act=act)
                out[i] = o.permute(0, 3, 1, 2)
        return out
