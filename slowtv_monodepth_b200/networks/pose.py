"""Relative pose (+ optional intrinsics) network — host-side mirror of `src/networks/pose.py` (reference)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .. import _lib as L_
from .. import functional as F_
from .encoders import create_encoder

__all__ = ['PoseNet']


def _block(cin: int, cout: int, k: int, pad: int = 0) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, k, 1, pad), nn.ReLU(inplace=True))


def _head(c: int, cout: int) -> nn.Sequential:
    """conv3x3+ReLU, conv3x3+ReLU, conv1x1 (the pooling / activation tail is applied functionally in `forward`)."""
    return nn.Sequential(_block(c, c, 3, 1), _block(c, c, 3, 1), nn.Conv2d(c, cout, 1))


class PoseNet(nn.Module):
    """Reference: src/networks/pose.py:14-135. Output contract identical: {'R','t': (b,2,3)[, 'fs','cs': (b,2)]}.
    Parameter names match (`squeeze.0`, `decoders.{pose,focal,offset}.{0.0,1.0,2}`)."""
    def __init__(self, enc_name: str = 'resnet18', learn_K: bool = False, pretrained: bool = False):
        super().__init__()
        self.enc_name, self.learn_K, self.pretrained = enc_name, learn_K, pretrained
        self.n_imgs, self.n_ch_dec, self.pose_eps = 2, 256, 0.01
        self.encoder = create_encoder(enc_name, in_chans=3*self.n_imgs, pretrained=pretrained)
        self.n_ch_enc = self.encoder.feature_info.channels()
        self.squeeze = _block(self.n_ch_enc[-1], self.n_ch_dec, 1)
        self.decoders = nn.ModuleDict({'pose': _head(self.n_ch_dec, 6*self.n_imgs)})
        if learn_K:
            self.decoders['focal'] = _head(self.n_ch_dec, 2)
            self.decoders['offset'] = _head(self.n_ch_dec, 2)

    @staticmethod
    def build_K(fs: Tensor, cs: Tensor) -> Tensor:
        """(b,2) focal lengths and principal points -> (b,4,4) intrinsics (pose.py:61-73)."""
        z, o = torch.zeros_like(fs[:, 0]), torch.ones_like(fs[:, 0])
        return torch.stack([fs[:, 0], z, cs[:, 0], z,
                            z, fs[:, 1], cs[:, 1], z,
                            z, z, o, z,
                            z, z, z, o], dim=-1).unflatten(-1, (4, 4))

    @staticmethod
    def _head_nhwc(head: nn.Sequential, feat: Tensor) -> Tensor:
        """conv3x3+ReLU, conv3x3+ReLU, conv1x1 on a channels-last tensor, then the spatial mean -> (B, cout)."""
        y = F_.conv2d_nhwc(feat, head[0][0].weight, head[0][0].bias, pad=1, act='relu')
        y = F_.conv2d_nhwc(y, head[1][0].weight, head[1][0].bias, pad=1, act='relu')
        return F_.conv2d_nhwc(y, head[2].weight, head[2].bias).mean(dim=(1, 2))

    def forward(self, x: Tensor) -> dict:
        if x.is_cuda: return self.forward_nhwc(x)
        return L_.host_path(self, x)

    def forward_nhwc(self, x: Tensor) -> dict:
        """Same network; every convolution is a libstv tcgen05 implicit GEMM on channels-last tensors."""
        feat = self.encoder(x)[-1].permute(0, 2, 3, 1)
        feat = F_.conv2d_nhwc(feat, self.squeeze[0].weight, self.squeeze[0].bias, act='relu')
        out = self.pose_eps*self._head_nhwc(self.decoders['pose'], feat).unflatten(-1, (self.n_imgs, 6))
        res = {'R': out[..., :3], 't': out[..., 3:]}
        if self.learn_K:
            res['fs'] = F.softplus(self._head_nhwc(self.decoders['focal'], feat))
            res['cs'] = torch.sigmoid(self._head_nhwc(self.decoders['offset'], feat))
        return res
