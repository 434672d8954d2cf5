"""Reference checkpoints -> B200 modules.

The reference saves Lightning checkpoints of `MonoDepthModule` (`api/train/train.py`; loaded again by
`api/quickstart/run.py:21-33`): `ckpt['state_dict']` holds the networks under `nets.<name>.<param>` (timm `FeatureListNet`
naming inside), `ckpt['hyper_parameters']['cfg']` the configuration the module was built from. The B200 modules keep the same
parameter names and shapes (`tests/test_plugin_cpu.py`, `tests/test_checkpoint_cpu.py`), so loading is a prefix strip plus a
strict `load_state_dict`; convolution filters are then re-laid channels-last by the flat optimiser buffer on first use.
"""
from __future__ import annotations

from pathlib import Path

import torch
import torch.nn as nn

__all__ = ['split_state_dict', 'load_nets', 'load_depth_net', 'save_nets']

PREFIX = 'nets.'


def split_state_dict(state_dict: dict) -> dict[str, dict]:
    """{'nets.depth.encoder.x': t, ...} -> {'depth': {'encoder.x': t}, ...}; keys outside `nets.` (losses, metrics) are ignored."""
    out: dict[str, dict] = {}
    for k, v in state_dict.items():
        if not k.startswith(PREFIX): continue
        name, _, rest = k[len(PREFIX):].partition('.')
        if rest: out.setdefault(name, {})[rest] = v
    return out


def _read(ckpt) -> dict:
    if isinstance(ckpt, (str, Path)): ckpt = torch.load(ckpt, map_location='cpu', weights_only=False)
    if not isinstance(ckpt, dict) or 'state_dict' not in ckpt: raise ValueError('Not a Lightning checkpoint: missing "state_dict".')
    return ckpt


def load_nets(nets: nn.ModuleDict, ckpt, strict: bool = True) -> list[str]:
    """Load every network of `nets` (e.g. `MonoDepthStep.nets`) that the checkpoint holds. Returns the names loaded.

    :raises KeyError: a network of `nets` is absent from the checkpoint (strict), as `load_state_dict` would for its keys."""
    parts = split_state_dict(_read(ckpt)['state_dict'])
    loaded = []
    for name, net in nets.items():
        if name not in parts:
            if strict: raise KeyError(f'Checkpoint has no weights for network "{name}" (found: {sorted(parts)}).')
            continue
        net.load_state_dict(parts[name], strict=strict)
        loaded.append(name)
    return loaded


def load_depth_net(ckpt, device=None) -> nn.Module:
    """The reference's quickstart (`api/quickstart/run.py:21-33`): build `DepthNet(**cfg['net']['depth'])` from the checkpoint's own
    hyper-parameters and load its weights. `pretrained` is forced off (no download; the weights come from the checkpoint)."""
    from .networks import DepthNet
    ckpt = _read(ckpt)
    try: cfg = dict(ckpt['hyper_parameters']['cfg']['net']['depth'])
    except KeyError as e: raise KeyError('Checkpoint carries no hyper_parameters.cfg.net.depth entry.') from e
    cfg['pretrained'] = False
    net = DepthNet(**cfg)
    net.load_state_dict(split_state_dict(ckpt['state_dict'])['depth'])
    for p in net.parameters(): p.requires_grad = False
    net.eval()
    return net.to(device) if device is not None else net


def save_nets(nets: nn.ModuleDict, path, cfg: dict | None = None) -> None:
    """Write a checkpoint the REFERENCE can read back (same key layout), e.g. to evaluate B200-trained weights with api/eval."""
    sd = {f'{PREFIX}{name}.{k}': v.detach().cpu() for name, net in nets.items() for k, v in net.state_dict().items()}
    torch.save({'state_dict': sd, 'hyper_parameters': {'cfg': cfg or {}}}, path)
