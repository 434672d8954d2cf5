"""ORACLE (test infrastructure — not shipped, not on the product path).

CPU restatement of the networks on the reference's hot path, in plain `torch.nn` (any dtype, any device, stock ATen ops):

  * the timm feature extractors the reference instantiates with `timm.create_model(name, features_only=True, ...)`
    (src/networks/depth.py:97, src/networks/pose.py:40). timm==0.6.12 (docker/environment.yml:287) is an un-vendored
    third-party dependency and is not installed here, so this is a restatement of its published ResNet / ConvNeXt
    definitions — PARITY UNPINNED at this boundary (no timm, no reference tests); module names follow timm's
    `FeatureListNet` so that checkpoints line up;
  * `DepthNet`, `MonodepthDecoder`, `PoseNet` (src/networks/depth.py:17-156, decoders/monodepth.py:15-89,
    decoders/utils.py:44-54, pose.py:14-135). These ARE pinned: `tests/test_oracle_nets.py` runs the reference's own
    classes (on top of `timm_create_model` below, via oracle/ref_shim.py) against these restatements with shared weights.

`timm_create_model` / `timm_create_optimizer_v2` are what oracle/ref_shim.py installs as the `timm` placeholder.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------------
# timm feature extractors
# ---------------------------------------------------------------------------------------------------------------------
class _Info:
    def __init__(self, ch, red): self.ch, self.red = ch, red
    def channels(self): return list(self.ch)
    def reduction(self): return list(self.red)


def _res_block(cin, cout, stride):
    m = nn.Module()
    m.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False); m.bn1 = nn.BatchNorm2d(cout)
    m.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False); m.bn2 = nn.BatchNorm2d(cout)
    if stride != 1 or cin != cout:
        m.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
    return m


class ResNetFeatures(nn.Module):
    """timm `resnet18/34` BasicBlock trunk; features after act1 and layer1..4."""
    def __init__(self, layers=(2, 2, 2, 2), in_chans=3):
        super().__init__()
        self.conv1 = nn.Conv2d(in_chans, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        widths, cin = (64, 128, 256, 512), 64
        for li, (w, nb) in enumerate(zip(widths, layers), start=1):
            blocks = []
            for bi in range(nb):
                blocks.append(_res_block(cin, w, 2 if (bi == 0 and li > 1) else 1))
                cin = w
            self.add_module(f'layer{li}', nn.ModuleList(blocks))
        self.feature_info = _Info((64, 64, 128, 256, 512), (2, 4, 8, 16, 32))
        for m in self.modules():
            if isinstance(m, nn.Conv2d): nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        for n, m in self.named_modules():
            if n.endswith('bn2'): nn.init.zeros_(m.weight)  # timm zero_init_last

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(x)))
        feats = [x]
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        for li in range(1, 5):
            for blk in getattr(self, f'layer{li}'):
                idt = blk.downsample(x) if hasattr(blk, 'downsample') else x
                out = torch.relu(blk.bn1(blk.conv1(x)))
                out = blk.bn2(blk.conv2(out))
                x = torch.relu(out + idt)
            feats.append(x)
        return feats


def _ln_channels(x, weight, bias):
    """LayerNorm over the channel axis of NCHW, eps=1e-6 (timm LayerNorm2d)."""
    mu = x.mean(1, keepdim=True)
    var = ((x - mu)**2).mean(1, keepdim=True)
    return (x - mu)/torch.sqrt(var + 1e-6)*weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


class _LN2d(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight, self.bias = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c))

    def forward(self, x): return _ln_channels(x, self.weight, self.bias)


class ConvNeXtFeatures(nn.Module):
    """timm `convnext_*`: stem (4x4/4 conv + LN2d), 4 stages of [LN2d + 2x2/2 conv] + blocks
    (dw7x7 -> LN -> Linear 4x -> GELU -> Linear -> gamma (1e-6) -> residual); features after each stage."""
    def __init__(self, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), in_chans=3):
        super().__init__()
        self.stem_0 = nn.Conv2d(in_chans, dims[0], 4, 4)
        self.stem_1 = _LN2d(dims[0])
        cin = dims[0]
        for si, (d, c) in enumerate(zip(depths, dims)):
            st = nn.Module()
            st.downsample = nn.Sequential(_LN2d(cin), nn.Conv2d(cin, c, 2, 2)) if si > 0 else nn.Identity()
            blocks = []
            for _ in range(d):
                b = nn.Module()
                b.gamma = nn.Parameter(torch.full((c,), 1e-6))
                b.conv_dw = nn.Conv2d(c, c, 7, padding=3, groups=c)
                b.norm = _LN2d(c)
                b.mlp = nn.Module(); b.mlp.fc1 = nn.Linear(c, 4*c); b.mlp.fc2 = nn.Linear(4*c, c)
                blocks.append(b)
            st.blocks = nn.ModuleList(blocks)
            self.add_module(f'stages_{si}', st)
            cin = c
        self.feature_info = _Info(tuple(dims), (4, 8, 16, 32))
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=.02); nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.stem_1(self.stem_0(x))
        feats = []
        for si in range(4):
            st = getattr(self, f'stages_{si}')
            x = st.downsample(x)
            for b in st.blocks:
                y = b.norm(b.conv_dw(x)).permute(0, 2, 3, 1)
                y = b.mlp.fc2(F.gelu(b.mlp.fc1(y)))*b.gamma
                x = x + y.permute(0, 3, 1, 2)
            feats.append(x)
        return feats


_SPECS = {
    'resnet18': lambda ic: ResNetFeatures((2, 2, 2, 2), ic),
    'resnet34': lambda ic: ResNetFeatures((3, 4, 6, 3), ic),
    'convnext_tiny': lambda ic: ConvNeXtFeatures((3, 3, 9, 3), (96, 192, 384, 768), ic),
    'convnext_small': lambda ic: ConvNeXtFeatures((3, 3, 27, 3), (96, 192, 384, 768), ic),
    'convnext_base': lambda ic: ConvNeXtFeatures((3, 3, 27, 3), (128, 256, 512, 1024), ic),
}


def timm_create_model(name, features_only=True, pretrained=False, in_chans=3, **kw):
    assert features_only, 'the reference only builds feature extractors'
    if pretrained: raise RuntimeError('no pretrained weights offline')
    return _SPECS[name](in_chans)


def timm_create_optimizer_v2(model_or_params, opt='adamw', lr=None, weight_decay=0., **kw):
    """timm's factory: for a module, biases and 1-D parameters are excluded from weight decay; 'adamw' -> torch.optim.AdamW."""
    assert opt == 'adamw', opt
    if isinstance(model_or_params, nn.Module):
        decay, no_decay = [], []
        for n, p in model_or_params.named_parameters():
            if not p.requires_grad: continue
            (no_decay if (p.ndim <= 1 or n.endswith('.bias')) else decay).append(p)
        groups = [{'params': no_decay, 'weight_decay': 0.}, {'params': decay, 'weight_decay': weight_decay}]
        return torch.optim.AdamW(groups, lr=lr, weight_decay=0., **kw)
    return torch.optim.AdamW(model_or_params, lr=lr, weight_decay=weight_decay, **kw)


# ---------------------------------------------------------------------------------------------------------------------
# Reference networks restated
# ---------------------------------------------------------------------------------------------------------------------
class MonodepthDecoder(nn.Module):
    """src/networks/decoders/monodepth.py:15-89 (+ conv3x3 reflect / conv_block ELU, decoders/utils.py:44-54)."""
    def __init__(self, num_ch_enc, enc_sc, out_sc=(0, 1, 2, 3), out_ch=1):
        super().__init__()
        self.enc_sc, self.out_sc, dec = list(enc_sc), list(out_sc), [16, 32, 64, 128, 256]
        conv = lambda i, o: nn.Conv2d(i, o, 3, padding=1, padding_mode='reflect')
        block = lambda i, o: nn.Sequential(OrderedDict(conv=conv(i, o), act=nn.ELU()))
        self.names, mods = [], []
        for i in range(4, -1, -1):
            self.names.append(f'upconv_{i}_0'); mods.append(block(num_ch_enc[-1] if i == 4 else dec[i + 1], dec[i]))
            extra = num_ch_enc[self.enc_sc.index(2**i)] if 2**i in self.enc_sc else 0
            self.names.append(f'upconv_{i}_1'); mods.append(block(dec[i] + extra, dec[i]))
        for i in self.out_sc:
            self.names.append(f'outconv_{i}'); mods.append(conv(dec[i], out_ch))
        self.decoder = nn.ModuleList(mods)

    def forward(self, feat):
        g = lambda n: self.decoder[self.names.index(n)]
        out, x = {}, feat[-1]
        for i in range(4, -1, -1):
            x = F.interpolate(g(f'upconv_{i}_0')(x), scale_factor=2, mode='nearest')
            if 2**i in self.enc_sc: x = torch.cat([x, feat[self.enc_sc.index(2**i)]], 1)
            x = g(f'upconv_{i}_1')(x)
            if i in self.out_sc: out[i] = torch.sigmoid(g(f'outconv_{i}')(x))
        return out


class DepthNet(nn.Module):
    """src/networks/depth.py:17-156 (monodepth decoder, no masks / virtual stereo)."""
    def __init__(self, enc_name='resnet18', out_scales=(0, 1, 2, 3)):
        super().__init__()
        self.out_scales = list(out_scales)
        self.encoder = timm_create_model(enc_name, features_only=True)
        self.decoders = nn.ModuleDict({'disp': MonodepthDecoder(self.encoder.feature_info.channels(),
                                                                 self.encoder.feature_info.reduction(), self.out_scales)})

    def forward(self, x):
        feat = self.encoder(x)
        return {'depth_feats': feat, 'disp': dict(sorted(self.decoders['disp'](feat).items()))}


class PoseNet(nn.Module):
    """src/networks/pose.py:14-135."""
    def __init__(self, enc_name='resnet18', learn_K=False):
        super().__init__()
        self.learn_K = learn_K
        self.encoder = timm_create_model(enc_name, features_only=True, in_chans=6)
        c = 256
        blk = lambda i, o, k, p=0: nn.Sequential(nn.Conv2d(i, o, k, 1, p), nn.ReLU())
        head = lambda o: nn.Sequential(blk(c, c, 3, 1), blk(c, c, 3, 1), nn.Conv2d(c, o, 1))
        self.squeeze = blk(self.encoder.feature_info.channels()[-1], c, 1)
        self.decoders = nn.ModuleDict({'pose': head(12)})
        if learn_K: self.decoders['focal'], self.decoders['offset'] = head(2), head(2)

    def forward(self, x):
        f = self.squeeze(self.encoder(x)[-1])
        o = 0.01*self.decoders['pose'](f).mean((2, 3)).view(-1, 2, 6)
        out = {'R': o[..., :3], 't': o[..., 3:]}
        if self.learn_K:
            out['fs'] = F.softplus(self.decoders['focal'](f).mean((2, 3)))
            out['cs'] = torch.sigmoid(self.decoders['offset'](f).mean((2, 3)))
        return out
