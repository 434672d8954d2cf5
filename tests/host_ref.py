"""TEST INFRASTRUCTURE: host (CPU) reference arithmetic for the product's modules.

The product package is CUDA-only and contains no host arithmetic: a module handed CPU tensors raises `StvError`. The CPU test-suite
still wants to check parameter naming / wiring / state-dict compatibility / the multi-process optimiser logic without a GPU, so
the plain-ATen implementations of the same modules live HERE and are registered into the package's (otherwise empty) hook table
`slowtv_monodepth_b200._lib._HOST_HOOKS` by the fixtures of tests/conftest.py (and by spawned worker processes explicitly)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _basic_block(m, x):
    y = F.relu(m.bn1(m.conv1(x)))
    y = m.bn2(m.conv2(y))
    sc = x if m.downsample is None else m.downsample(x)
    return F.relu(y + sc)


def _resnet(m, x):
    f0 = F.relu(m.bn1(m.conv1(x)))
    x = F.max_pool2d(f0, 3, 2, 1)
    feats = [f0]
    for i in range(1, 5):
        x = getattr(m, f'layer{i}')(x)
        feats.append(x)
    return feats


def _ln2d(m, x):
    return F.layer_norm(x.permute(0, 2, 3, 1), m.normalized_shape, m.weight, m.bias, m.eps).permute(0, 3, 1, 2)


def _mlp(m, x): return m.fc2(F.gelu(m.fc1(x)))


def _cnx_block(m, x):
    y = m.conv_dw(x).permute(0, 2, 3, 1)
    y = m.mlp(m.norm(y))*m.gamma
    return x + y.permute(0, 3, 1, 2)


def _cnx_stage(m, x): return m.blocks(m.downsample(x))


def _convnext(m, x):
    x = m.stem_1(m.stem_0(x))
    feats = []
    for i in range(4):
        x = getattr(m, f'stages_{i}')(x)
        feats.append(x)
    return feats


def _conv_block(m, x): return F.elu(m.conv(x))


def _decoder(m, feat):
    from slowtv_monodepth_b200.networks.decoder import _ACT
    out, act = {}, _ACT[m.out_act]
    x = feat[-1]
    for i in range(4, -1, -1):
        x = m.layer(f'upconv_{i}_0')(x)
        x = F.interpolate(x, scale_factor=2, mode=m.upsample_mode)
        if m.use_skip and 2**i in m.enc_sc: x = torch.cat([x, feat[m.enc_sc.index(2**i)]], dim=1)
        x = m.layer(f'upconv_{i}_1')(x)
        if i in m.out_sc: out[i] = act(m.layer(f'outconv_{i}')(x)).contiguous()
    return out


def _pose(m, x):
    feat = m.squeeze(m.encoder(x)[-1])
    out = m.pose_eps*m.decoders['pose'](feat).mean(dim=(2, 3)).unflatten(-1, (m.n_imgs, 6))
    res = {'R': out[..., :3], 't': out[..., 3:]}
    if m.learn_K:
        res['fs'] = F.softplus(m.decoders['focal'](feat).mean(dim=(2, 3)))
        res['cs'] = torch.sigmoid(m.decoders['offset'](feat).mean(dim=(2, 3)))
    return res


@torch.no_grad()
def _adamw(opt, grad_scale: float = 1.0):
    """FlatAdamW's update rule on host tensors (the gloo tests of the multi-process logic)."""
    b1, b2 = opt.betas
    g = opt.grad*(grad_scale/opt.world)
    for start, n, n_dec in opt.buckets: opt.flat[start:start + n_dec].mul_(1 - opt.lr*opt.weight_decay)
    opt.exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    opt.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1**opt.step_count, 1 - b2**opt.step_count
    opt.flat.addcdiv_(opt.exp_avg, opt.exp_avg_sq.sqrt()/bc2**0.5 + opt.eps, value=-opt.lr/bc1)


def install() -> None:
    from slowtv_monodepth_b200 import _lib
    from slowtv_monodepth_b200.networks import decoder, encoders, pose
    from slowtv_monodepth_b200.optim import FlatAdamW
    _lib._HOST_HOOKS.update({
        encoders.BasicBlock: _basic_block, encoders.ResNetEncoder: _resnet, encoders.LayerNorm2d: _ln2d, encoders.Mlp: _mlp,
        encoders.ConvNeXtBlock: _cnx_block, encoders.ConvNeXtStage: _cnx_stage, encoders.ConvNeXtEncoder: _convnext,
        decoder._ConvBlock: _conv_block, decoder.MonodepthDecoder: _decoder, pose.PoseNet: _pose, FlatAdamW: _adamw,
    })


def uninstall() -> None:
    from slowtv_monodepth_b200 import _lib
    _lib._HOST_HOOKS.clear()
