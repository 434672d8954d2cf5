"""Feature-pyramid encoders with timm's `features_only=True` contract (the reference builds them with
`timm.create_model(enc_name, features_only=True, ...)`, src/networks/depth.py:97, src/networks/pose.py:40).

timm (pinned 0.6.12, docker/environment.yml:287) is a third-party dependency that is not part of the reference tree, so
the arithmetic is restated here from its published architectures, keeping timm's module / parameter names so that a
reference checkpoint's `nets.*.encoder.*` keys line up:

  resnet18        conv1, bn1, act1, maxpool, layer{1..4}.{0,1}.{conv1,bn1,conv2,bn2,downsample.{0,1}}
                  features = outputs of act1, layer1..4; channels [64,64,128,256,512], reductions [2,4,8,16,32]
  convnext_{tiny,small,base}
                  stem_0 (4x4 s4 conv), stem_1 (LayerNorm2d), stages_{0..3}.{downsample.{0,1}, blocks.N.{conv_dw,norm,
                  mlp.fc1,mlp.fc2,gamma}};  features = outputs of stages_0..3; reductions [4,8,16,32]

ConvNeXt runs channels-last end to end (its LayerNorm / Linear layers are over the channel axis); the pointwise MLP is
the GEMM-shaped part that goes to the tcgen05 kernels (see ops_gemm.py), the depthwise 7x7 + LayerNorm is HBM-bound.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .. import _lib as L_
from .. import functional as F_

__all__ = ['create_encoder', 'ResNetEncoder', 'ConvNeXtEncoder', 'FeatureInfo']


def _stem_conv_nhwc(x: Tensor, conv: nn.Conv2d) -> Tensor:
    """Stem convolution of an encoder on the (N,C,H,W) image batch -> channels-last (N,P,Q,Cout).

    The 3 / 6 input channels are too narrow for the tensor-core loaders (a TMA im2col box is 32 channels), so both stems are
    rewritten, arithmetically exactly, as stride-1 problems over a space-to-depth view of the input:
      * ConvNeXt `stem_0` (k x k, stride k, no padding — a patchify): one GEMM over (k*k*C)-long pixel-patch rows;
      * ResNet `conv1` (7x7, stride 2, padding 3): the filter is zero-extended to 8x8 (one leading zero row / column), the
        input zero-padded (4 before, 2 after) and 2x2 space-to-depth'ed with C zero-padded to 8, which turns it into a 4x4,
        stride-1, unpadded convolution over 32 channels.
    The filter rearrangements are tiny autograd ops on the parameter, so its shape / name / gradient stay the checkpoint's."""
    N, Cc, H, W = x.shape
    k, st, pd, Cout = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.out_channels
    if k == st and pd == 0 and H % k == 0 and W % k == 0 and (k*k*Cc) % 4 == 0:
        rows = x.view(N, Cc, H//k, k, W//k, k).permute(0, 2, 4, 3, 5, 1).reshape(-1, k*k*Cc)      # (r, s, c) within a patch
        wmat = conv.weight.permute(0, 2, 3, 1).reshape(Cout, k*k*Cc)
        return F_.linear(rows, wmat, conv.bias).view(N, H//k, W//k, Cout)
    if k == 7 and st == 2 and pd == 3 and H % 2 == 0 and W % 2 == 0 and Cc <= 8:
        xp = F.pad(x.permute(0, 2, 3, 1), (0, 8 - Cc, 4, 2, 4, 2))                               # (N, H+6, W+6, 8)
        x2 = xp.view(N, (H + 6)//2, 2, (W + 6)//2, 2, 8).permute(0, 1, 3, 2, 4, 5).reshape(N, (H + 6)//2, (W + 6)//2, 32)
        w8 = F.pad(conv.weight, (1, 0, 1, 0, 0, 8 - Cc))                                         # (Cout, 8, 8, 8): c, r', s'
        w2 = w8.view(Cout, 8, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4).reshape(Cout, 32, 4, 4)      # channel = (uy, ux, c); taps (ty, tx)
        return F_.conv2d_nhwc(x2, w2, conv.bias)
    pad = (-Cc) % 4
    x4 = F.pad(x.permute(0, 2, 3, 1), (0, pad)) if pad else x.permute(0, 2, 3, 1).contiguous()
    w = F.pad(conv.weight, (0, 0, 0, 0, 0, pad)) if pad else conv.weight
    return F_.conv2d_nhwc(x4, w, conv.bias, stride=st, pad=pd)


def _bn_nhwc(x: Tensor, bn: nn.BatchNorm2d, relu: bool = False, res: Tensor | None = None) -> Tensor:
    """[relu](BatchNorm(x) [+ res]) on a channels-last tensor. Training mode (per-GPU batch statistics, as the reference): one
    libstv launch set with the residual add and ReLU fused; evaluation mode (running statistics) is a plain affine map."""
    if bn.training:
        return F_.batch_norm_nhwc(x, bn.weight, bn.bias, res=res, relu=relu, run_mean=bn.running_mean, run_var=bn.running_var,
                                  eps=bn.eps, momentum=bn.momentum)
    scale = bn.weight*torch.rsqrt(bn.running_var + bn.eps)
    y = x*scale + (bn.bias - bn.running_mean*scale)
    if res is not None: y = y + res
    return F.relu(y) if relu else y


class FeatureInfo:
    """Minimal stand-in for timm's `feature_info` (only what the reference reads: depth.py:98, pose.py:41)."""
    def __init__(self, channels: list[int], reduction: list[int]):
        self._c, self._r = list(channels), list(reduction)

    def channels(self) -> list[int]: return list(self._c)
    def reduction(self) -> list[int]: return list(self._r)


# ---------------------------------------------------------------------------------------------------------------------
# ResNet (BasicBlock)
# ---------------------------------------------------------------------------------------------------------------------
class BasicBlock(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda: return self.forward_nhwc(x)
        return L_.host_path(self, x)

    def forward_nhwc(self, x: Tensor) -> Tensor:
        """x (N,H,W,C) channels-last; convolutions are libstv tcgen05 implicit GEMMs."""
        s = self.conv1.stride[0]
        c1 = F_.conv2d_nhwc(x, self.conv1.weight, None, stride=s, pad=1)
        self.first_node = c1.grad_fn   # created before the shortcut's node: the block's last node to run in backward
        y = _bn_nhwc(c1, self.bn1, relu=True)
        sc = x if self.downsample is None else _bn_nhwc(F_.conv2d_nhwc(x, self.downsample[0].weight, None, stride=s), self.downsample[1])
        return _bn_nhwc(F_.conv2d_nhwc(y, self.conv2.weight, None, pad=1), self.bn2, relu=True, res=sc)


class ResNetEncoder(nn.Module):
    def __init__(self, layers=(2, 2, 2, 2), in_chans: int = 3):
        super().__init__()
        self.conv1 = nn.Conv2d(in_chans, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        chans = [64, 128, 256, 512]
        cin = 64
        for i, (c, nblk) in enumerate(zip(chans, layers)):
            blocks = [BasicBlock(cin if j == 0 else c, c, (1 if i == 0 else 2) if j == 0 else 1) for j in range(nblk)]
            setattr(self, f'layer{i + 1}', nn.Sequential(*blocks))
            cin = c
        self.feature_info = FeatureInfo([64, 64, 128, 256, 512], [2, 4, 8, 16, 32])
        self.reset_parameters()

    def reset_parameters(self) -> None:
        # timm resnet init: kaiming-normal (fan_out, relu) convs; BN weight 1 / bias 0; zero_init_last on bn2.
        for m in self.modules():
            if isinstance(m, nn.Conv2d): nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        for m in self.modules():
            if isinstance(m, BasicBlock): nn.init.zeros_(m.bn2.weight)

    def forward(self, x: Tensor) -> list[Tensor]:
        if x.is_cuda:
            if self.training:   # nn.BatchNorm2d counts its training forwards: one multi-tensor launch for all layers of the encoder
                if getattr(self, '_nbt', None) is None or (self._nbt and self._nbt[0].device != x.device):
                    self._nbt = [m.num_batches_tracked for m in self.modules() if isinstance(m, nn.BatchNorm2d) and m.num_batches_tracked is not None]
                if self._nbt: torch._foreach_add_(self._nbt, 1)
            # `marks`: autograd node of the first operation of each part; when it has run in backward, every gradient of that part
            # (and of everything after it in the forward order) is final — FlatAdamW starts that bucket's all-reduce from it.
            y = _stem_conv_nhwc(x, self.conv1)
            self.marks = {'stem': y.grad_fn}
            f0 = _bn_nhwc(y, self.bn1, relu=True)
            x = F_.maxpool3x3s2(f0)
            feats = [f0]
            for i in range(1, 5):
                for j, blk in enumerate(getattr(self, f'layer{i}')):
                    x = blk.forward_nhwc(x)
                    if j == 0: self.marks[f'layer{i}'] = blk.first_node
                feats.append(x)
            return [f.permute(0, 3, 1, 2) for f in feats]  # (N,C,H,W) views of the channels-last buffers
        return L_.host_path(self, x)


# ---------------------------------------------------------------------------------------------------------------------
# ConvNeXt
# ---------------------------------------------------------------------------------------------------------------------
class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over the channel axis of an NCHW tensor (timm `LayerNorm2d`, eps 1e-6)."""
    def __init__(self, c: int): super().__init__(c, eps=1e-6)

    def forward(self, x: Tensor) -> Tensor:
        return L_.host_path(self, x)   # on the device the owning stage calls functional.layer_norm on the channels-last tensor


class Mlp(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.fc1 = nn.Linear(c, 4*c)
        self.fc2 = nn.Linear(4*c, c)

    def forward(self, x: Tensor) -> Tensor:
        return L_.host_path(self, x)   # on the device the owning block calls functional.convnext_mlp


class ConvNeXtBlock(nn.Module):
    """x + gamma * fc2(GELU(fc1(LN(dwconv7x7(x)))))  (timm `ConvNeXtBlock`, ls_init_value=1e-6, no drop-path at train default)."""
    def __init__(self, c: int):
        super().__init__()
        self.conv_dw = nn.Conv2d(c, c, 7, 1, 3, groups=c)
        self.norm = nn.LayerNorm(c, eps=1e-6)
        self.mlp = Mlp(c)
        self.gamma = nn.Parameter(1e-6*torch.ones(c))

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda: return self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
        return L_.host_path(self, x)

    def forward_nhwc(self, xl: Tensor) -> Tensor:
        """xl (N,H,W,C) channels-last. libstv kernels: depthwise 7x7 (fwd / dgrad / wgrad), LayerNorm (fwd / bwd) and the
        pointwise MLP as tcgen05 TF32 GEMMs with bias+GELU / bias+layer-scale+residual epilogues."""
        # `link` carries the residual-branch gradient from the MLP's backward to the depthwise data-gradient kernel, which adds
        # it in its epilogue (xl feeds both): one pass less over the activation than autograd's own accumulation.
        link = {} if xl.requires_grad and xl.is_contiguous() else None
        y = F_.dwconv7(xl, self.conv_dw.weight, self.conv_dw.bias, link)
        y = F_.layer_norm(y, self.norm.weight, self.norm.bias, self.norm.eps)
        c = xl.shape[-1]
        out = F_.convnext_mlp(y.view(-1, c), xl.reshape(-1, c), self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight,
                              self.mlp.fc2.bias, self.gamma, link)
        return out.view(xl.shape)


class ConvNeXtStage(nn.Module):
    def __init__(self, cin: int, cout: int, depth: int, stride: int):
        super().__init__()
        self.downsample = nn.Sequential(LayerNorm2d(cin), nn.Conv2d(cin, cout, 2, 2)) if stride > 1 else nn.Identity()
        self.blocks = nn.Sequential(*[ConvNeXtBlock(cout) for _ in range(depth)])

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda: return self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
        return L_.host_path(self, x)

    def forward_nhwc(self, x: Tensor, want_mark: bool = False):
        mark = None
        if not isinstance(self.downsample, nn.Identity):
            ln, conv = self.downsample[0], self.downsample[1]
            x = F_.layer_norm(x, ln.weight, ln.bias, ln.eps)
            mark = x.grad_fn
            x = F_.conv2d_nhwc(x, conv.weight, conv.bias, stride=2)
        for blk in self.blocks: x = blk.forward_nhwc(x)
        return (x, mark) if want_mark else x   # stage 0 has no first op of its own (mark None): it completes with the stem


class ConvNeXtEncoder(nn.Module):
    def __init__(self, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), in_chans: int = 3):
        super().__init__()
        self.stem_0 = nn.Conv2d(in_chans, dims[0], 4, 4)
        self.stem_1 = LayerNorm2d(dims[0])
        cin = dims[0]
        for i, (d, c) in enumerate(zip(depths, dims)):
            setattr(self, f'stages_{i}', ConvNeXtStage(cin, c, d, 1 if i == 0 else 2))
            cin = c
        self.feature_info = FeatureInfo(list(dims), [4, 8, 16, 32])
        self.reset_parameters()

    def reset_parameters(self) -> None:
        for m in self.modules():  # timm convnext `_init_weights`: trunc_normal(std=.02) weights, zero biases
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None: nn.init.zeros_(m.bias)

    def forward(self, x: Tensor) -> list[Tensor]:
        if x.is_cuda:
            x = _stem_conv_nhwc(x, self.stem_0)
            self.marks = {'stem': x.grad_fn}   # see ResNetEncoder.forward
            x = F_.layer_norm(x, self.stem_1.weight, self.stem_1.bias, self.stem_1.eps)
            feats = []
            for i in range(4):
                st = getattr(self, f'stages_{i}')
                x, self.marks[f'stages_{i}'] = st.forward_nhwc(x, want_mark=True)
                feats.append(x.permute(0, 3, 1, 2))  # (N,C,H,W) view of the channels-last buffer
            return feats
        return L_.host_path(self, x)


_RESNETS = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3)}
_CONVNEXTS = {
    'convnext_tiny': ((3, 3, 9, 3), (96, 192, 384, 768)),
    'convnext_small': ((3, 3, 27, 3), (96, 192, 384, 768)),
    'convnext_base': ((3, 3, 27, 3), (128, 256, 512, 1024)),
}


def create_encoder(name: str, in_chans: int = 3, pretrained: bool = False) -> nn.Module:
    """`timm.create_model(name, features_only=True, in_chans=..., pretrained=...)` for the encoders the KBR configs use."""
    if pretrained:
        raise RuntimeError('ImageNet-pretrained weights are not available offline; load a checkpoint with load_state_dict instead.')
    if name in _RESNETS: return ResNetEncoder(_RESNETS[name], in_chans)
    if name in _CONVNEXTS: return ConvNeXtEncoder(*_CONVNEXTS[name], in_chans)
    raise KeyError(f'Unknown encoder "{name}". Available: {sorted(_RESNETS) + sorted(_CONVNEXTS)}')
