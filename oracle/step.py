"""ORACLE (test infrastructure — not shipped, not on the product path).

CPU restatement of one full training step of the reference: `MonoDepthModule.step` (src/core/trainer.py:115-190:
forward :192-278, forward_postprocess :280-348, forward_loss :350-472 for {img_recon, disp_smooth}) followed by
Lightning's `loss.backward()` and the AdamW step built by `parsers.get_opt` (src/tools/parsers.py:205-243).
Used by the parity tests and as the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import loss as OL
from . import nets as ON


class OracleTrainer(nn.Module):
    def __init__(self, depth_enc='convnext_tiny', pose_enc='resnet18', learn_K=False, lr=1e-4, weight_decay=1e-3,
                 min_depth=0.1, max_depth=100., always_fwd_pose=False, w_recon=1., w_smooth=1e-3):
        super().__init__()
        self.nets = nn.ModuleDict({'depth': ON.DepthNet(depth_enc), 'pose': ON.PoseNet(pose_enc, learn_K)})
        self.min_depth, self.max_depth, self.always_fwd_pose = min_depth, max_depth, always_fwd_pose
        self.w_recon, self.w_smooth = w_recon, w_smooth
        self.opt = ON.timm_create_optimizer_v2(self.nets, 'adamw', lr=lr, weight_decay=weight_decay)

    def forward_nets(self, x):
        fwd = dict(self.nets['depth'](x['imgs']))
        idxs = [int(i) for i in x['supp_idxs']]
        inv = lambda i: self.always_fwd_pose and i < 0
        pairs = torch.stack([torch.cat([s, x['imgs']] if inv(i) else [x['imgs'], s], 1) for i, s in zip(idxs, x['supp_imgs'])])
        sh = pairs.shape[:2]
        out = self.nets['pose'](pairs.flatten(0, 1))
        Ts = OL.T_from_AAt(out['R'][:, 0], out['t'][:, 0]).unflatten(0, sh)
        fwd['Ts'] = torch.stack([torch.linalg.inv(T) if inv(i) else T for i, T in zip(idxs, Ts)])
        if 'fs' in out:
            K = OL.build_K(out['fs'], out['cs']).unflatten(0, sh)[0]
            fwd['K'] = OL.resize_K(K, x['imgs'].shape[-2:])
        return fwd

    def loss(self, batch, noise=None, forced_sel=None):
        x, y, _ = batch
        fwd = self.forward_nets(x)
        disps = [fwd['disp'][s] for s in sorted(fwd['disp'])]
        loss, out = OL.loss_stack(disps, y['imgs'], y['supp_imgs'], fwd['Ts'], fwd.get('K', y['K']), self.min_depth,
                                  self.max_depth, self.w_recon, self.w_smooth, noise=noise, forced_sel=forced_sel)
        return loss, out, fwd

    def train_step(self, batch, noise=None):
        self.opt.zero_grad(set_to_none=True)
        loss, _, _ = self.loss(batch, noise)
        loss.backward()
        self.opt.step()
        return loss.detach()
