"""The per-step training loop — host-side mirror of `MonoDepthModule` (src/core/trainer.py, reference) without Lightning.

`MonoDepthStep.step(batch)` follows the reference's `step` -> `forward` -> `forward_postprocess` -> `forward_loss`
(trainer.py:115-190, 192-278, 280-348, 350-472) for the KBR loss set {img_recon, disp_smooth}, with the differences that
matter on a B200:
  * no device syncs inside the step (the reference's timers sync 18x per step, trainer.py:69, and its f-string / comparison
    use of the device tensor `supp_idxs` adds ~5 more per support frame, trainer.py:242-253);
  * upsample + disp->depth, the whole view-synthesis loss and the smoothness term are libstv kernels;
  * the pose network sees all n*b image pairs in one batch (as the reference does, trainer.py:243-249).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as F_
from . import geometry as G
from . import handlers as H
from .losses import ReconstructionLoss
from .networks import DepthNet, PoseNet
from .regularizers import SmoothReg

__all__ = ['MonoDepthStep', 'GraphedTrainStep', 'ShapeCachedTrainStep', 'StepSummary', 'summarize', 'default_cfg', 'gradient_buckets',
           'register_bucket_hooks', 'clear_bucket_marks']

NET_REG = {'depth': DepthNet, 'pose': PoseNet}
LOSS_REG = {'img_recon': ReconstructionLoss, 'disp_smooth': SmoothReg}


def default_cfg(depth_enc: str = 'convnext_tiny', pose_enc: str = 'resnet18', learn_K: bool = False) -> dict:
    """The KBR-style configuration of BASELINE.json config 3 (cfg/abl_learn_K/default.yaml + cfg/kbr/default.yaml:18-28,105-125)."""
    return {
        'net': {'depth': {'enc_name': depth_enc, 'pretrained': False, 'dec_name': 'monodepth', 'out_scales': [0, 1, 2, 3]},
                'pose': {'enc_name': pose_enc, 'pretrained': False, 'learn_K': learn_K}},
        'loss': {'img_recon': {'weight': 1, 'loss_name': 'ssim', 'use_min': True, 'use_automask': True},
                 'disp_smooth': {'weight': 0.001, 'use_edges': True}},
        'optimizer': {'type': 'adamw', 'lr': 1e-4, 'weight_decay': 1e-3},
        'trainer': {'min_depth': 0.1, 'max_depth': 100, 'always_fwd_pose': False},
    }


class MonoDepthStep(nn.Module):
    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        nets = {}
        for k, kw in cfg['net'].items():
            if kw is None: continue
            if k not in NET_REG: raise KeyError(f'Unrecognized key: {k}.')
            nets[k] = NET_REG[k](**kw)
        self.nets = nn.ModuleDict(nets)
        self.losses, self.weights = {}, {}
        for k, kw in cfg['loss'].items():
            if kw is None: continue
            if k not in LOSS_REG: raise ValueError(f'Missing loss key: "{k}"')
            kw = dict(kw)
            self.weights[k] = float(kw.pop('weight', 1))
            self.losses[k] = LOSS_REG[k](**kw)
        tr = cfg.get('trainer', {})
        self.min_depth, self.max_depth = tr.get('min_depth'), tr.get('max_depth')
        self.always_fwd_pose = tr.get('always_fwd_pose', True)
        self.scales = self.nets['depth'].out_scales

    # -- trainer.py:192-278 ------------------------------------------------------------------------------------------
    def forward(self, x: dict) -> dict:
        fwd = {}
        idxs = [int(i) for i in x['supp_idxs']]  # host ints: no device sync
        # The pose network is independent of the depth network: it runs on its own stream, forward and backward, beside it
        # (functional.branch_stream); everything it produces is joined before the loss reads it.
        br = None
        if 'pose' in self.nets and any(i != 0 for i in idxs):   # stereo-only batches have no motion to predict
            inv = lambda i: self.always_fwd_pose and i < 0
            with F_.branch_stream('pose', [x['imgs'], x['supp_imgs']]) as br:
                pairs = torch.stack([torch.cat([s, x['imgs']] if inv(i) else [x['imgs'], s], dim=1)
                                     for i, s in zip(idxs, x['supp_imgs']) if i != 0])  # (n, b, 6, h, w)
                sh = pairs.shape[:2]
                out = self.nets['pose'](pairs.flatten(0, 1))
                Ts = G.T_from_AAt(aa=out['R'][:, 0], t=out['t'][:, 0]).unflatten(0, sh)
                for i, T in zip([i for i in idxs if i != 0], Ts):
                    fwd[f'T_{i}'] = F_.inv4x4(T) if inv(i) else T
                if 'fs' in out:
                    fwd['fs'], fwd['cs'] = out['fs'].unflatten(0, sh), out['cs'].unflatten(0, sh)
                    K = PoseNet.build_K(out['fs'], out['cs']).unflatten(0, sh)[0]  # first support frame only (trainer.py:259)
                    fwd['K'] = G.resize_K(K, x['imgs'].shape[-2:])
            pose_out = [v for v in fwd.values() if torch.is_tensor(v)]
        fwd |= self.nets['depth'](x['imgs'])
        if br is not None: br.join(pose_out)
        fwd['_idxs'] = idxs
        hook = getattr(self, 'bucket_hook', None)
        if hook is not None and torch.is_grad_enabled(): register_bucket_hooks(self.nets, hook)
        clear_bucket_marks(self.nets)
        return fwd

    # -- trainer.py:280-348 ------------------------------------------------------------------------------------------
    def forward_postprocess(self, fwd: dict, x: dict, y: dict, want_up: bool = True) -> dict:
        """`disp_up` / `depth_up` (trainer.py:316-321) are produced for the logging statistics and for callers of the reference's
        handler signature; the loss kernel itself starts from the low-resolution disparities (the maps are tagged with their
        source, `functional.disp_to_depth`), so nothing is differentiated through these up-sampled copies."""
        if want_up:
            size = x['imgs'].shape[-2:]
            up = {s: G.upsample_to_depth(d, size, self.min_depth, self.max_depth) for s, d in fwd['disp'].items()}
            fwd['disp_up'] = {s: v[0] for s, v in up.items()}
            fwd['depth_up'] = {s: v[1] for s, v in up.items()}
        # stereo support frames (index 0) take the calibrated baseline of the batch, temporal ones the predicted motion (trainer.py:347)
        Ts = []
        for i in fwd['_idxs']:
            if i == 0:
                if 'T_stereo' not in y: raise KeyError('Support index 0 (stereo pair) needs the baseline transform y["T_stereo"].')
                Ts.append(y['T_stereo'])
            elif f'T_{i}' not in fwd: raise KeyError(f'No pose for support index {i}: temporal support frames need a "pose" network.')
            else: Ts.append(fwd[f'T_{i}'])
        fwd['Ts'] = torch.stack(Ts)
        return fwd

    # -- trainer.py:350-472 ------------------------------------------------------------------------------------------
    def forward_loss(self, fwd: dict, x: dict, y: dict, want_maps: bool = False, noise: Tensor | None = None):
        loss, loss_dict = 0., {}
        for k, crit in self.losses.items():
            if k == 'img_recon':
                l, ld = H.image_recon(crit, None, depths=fwd.get('depth_up'), masks=None, imgs=y['imgs'], supp_imgs=y['supp_imgs'],
                                      Ts=fwd['Ts'], Ks=fwd.get('K', y['K']), want_warp=want_maps, noise=noise,
                                      disps=fwd['disp'], depth_range=(self.min_depth, self.max_depth))
            elif k == 'disp_smooth':
                l, ld = H.disp_smooth(crit, fwd['disp'], y['imgs'], want_maps=want_maps)
            else:
                raise ValueError(f'Missing loss key: "{k}"')
            loss = loss + self.weights[k]*l
            loss_dict[f'loss_{k}'] = l
            loss_dict.update(ld)
        return loss, loss_dict

    # -- trainer.py:115-190 ------------------------------------------------------------------------------------------
    def step(self, batch, mode: str = 'train', want_maps: bool = False, noise: Tensor | None = None, want_up: bool | None = None):
        """`noise`: explicit auto-mask tie-break noise (S*b,1,H,W) replacing the per-step draw (parity tests).
        `want_up`: also materialise the up-sampled `disp_up` / `depth_up` maps (logging statistics, `summarize`); by default only
        outside plain training steps — the loss does not need them."""
        x, y, m = batch
        fwd = self.forward(x)
        fwd = self.forward_postprocess(fwd, x, y, want_up=(want_maps or mode != 'train') if want_up is None else want_up)
        loss, loss_dict = self.forward_loss(fwd, x, y, want_maps=want_maps or mode != 'train', noise=noise)
        return loss, loss_dict, fwd


def gradient_buckets(nets) -> list[str]:
    """Parameter-name prefixes of `nets` (a ModuleDict with 'depth' and optionally 'pose') in the order their gradients become
    FINAL during backward: the pose network (built last in the forward, so autograd runs it first), the depth decoder, then the
    depth encoder from its deepest stage to the stem. `FlatAdamW(nets, buckets=...)` lays the flat buffers out in this order."""
    out = []
    if 'pose' in nets: out.append('pose.')
    out.append('depth.decoders.')
    enc = nets['depth'].encoder
    parts = [n for n, _ in enc.named_children() if n.startswith(('stages_', 'layer'))]
    out += [f'depth.encoder.{n}.' for n in reversed(parts[1:])]   # the first stage completes with the stem: last (implicit) bucket
    return out


def register_bucket_hooks(nets, bucket_ready) -> None:
    """After a forward pass: attach `bucket_ready(j)` to the autograd node whose completion makes bucket j of `gradient_buckets`
    final (the first operation of the corresponding part, recorded by the networks as `marks` / `first_node`)."""
    nodes = []
    if 'pose' in nets: nodes.append(getattr(nets['pose'].encoder, 'marks', {}).get('stem'))
    dec = nets['depth'].decoders['disp']
    nodes.append(getattr(dec, 'first_node', None))
    enc = nets['depth'].encoder
    parts = [n for n, _ in enc.named_children() if n.startswith(('stages_', 'layer'))]
    nodes += [getattr(enc, 'marks', {}).get(n) for n in reversed(parts[1:])]
    for j, node in enumerate(nodes):
        if node is not None: node.register_hook(lambda *a, j=j: bucket_ready(j))


def clear_bucket_marks(nets) -> None:
    """Drop the autograd nodes the networks recorded during the forward pass. A module attribute holding a `grad_fn` keeps the whole
    autograd graph of that step alive into the next one — including its AccumulateGrad nodes, which stay bound to the stream they
    were created on and then invalidate a CUDA-graph capture running on another stream."""
    for m in nets.modules():
        if getattr(m, 'marks', None): m.marks = {}
        if getattr(m, 'first_node', None) is not None: m.first_node = None


class StepSummary:
    """Names + ONE device vector of the step's logging scalars; `to_host()` is the only synchronisation."""
    def __init__(self, names: list[str], values: Tensor): self.names, self.values = names, values

    def to_host(self) -> dict[str, float]:
        return dict(zip(self.names, self.values.detach().cpu().tolist()))


@torch.no_grad()
def summarize(fwd: dict) -> StepSummary:
    """The reference's `summarize_depth` / `summarize_pose` / `summarize_K` (src/core/trainer.py:486-529, same keys) without their
    ~28 `.item()` synchronisations: the eight full-resolution maps go through one multi-tensor kernel (stv_mean_std), the
    pose / intrinsics statistics are a handful of tiny device ops, and everything lands in one vector."""
    names, parts = [], []
    maps, map_names = [], []
    for key in ('disp', 'depth'):
        for k, v in fwd.get(f'{key}_up', {}).items():
            maps.append(v); map_names += [f'{key}_mean_{k}', f'{key}_std_{k}']
    if maps:
        names += map_names; parts.append(F_.mean_std(maps).flatten())
    for k, v in fwd.items():
        if not (isinstance(k, str) and k.startswith('T_')): continue
        ts = v[..., :3, 3].pow(2).sum(-1).sqrt()
        tr = v[..., :3, :3].diagonal(dim1=-2, dim2=-1).mean(dim=-1)
        names += [f'{k}_t_mean', f'{k}_t_std', f'{k}_R_mean', f'{k}_R_std']
        parts.append(torch.stack([ts.mean(), ts.std(), tr.mean(), tr.std()]))
    if 'K' in fwd and 'fs' in fwd:
        names += ['fx', 'fy', 'cx', 'cy']
        parts.append(torch.stack([fwd['fs'][..., 0].mean(), fwd['fs'][..., 1].mean(), fwd['cs'][..., 0].mean(), fwd['cs'][..., 1].mean()]))
    vals = torch.cat([p.float() for p in parts]) if parts else torch.empty(0)
    return StepSummary(names, vals)


class GraphedTrainStep:
    """One training step (networks -> losses -> backward, + the overlapped gradient all-reduce when world > 1) captured ONCE as a
    CUDA graph and replayed every step.

    The step is ~1 800 kernel launches from ~1 000 Python-level operations; enqueueing them from the host takes longer than the
    GPU needs to execute them, so the eager loop is launch-bound. Capturing is possible because nothing on the path synchronises
    or allocates outside torch's graph-private pool: libstv entry points only enqueue kernels / memsets on the stream they are
    given (tensor maps are host-encoded kernel arguments), the matrix inverses are libstv kernels, not ATen's batched LU, and
    the auto-mask tie-break noise is seeded by a device-side counter that the loss kernel itself advances (fresh noise per replay,
    like the reference's per-step randn_like, reconstruction.py:72).

    Outside the graph: the gradient memset (so that `accumulate` > 1 micro-batches can add up before one optimiser step, Lightning's
    accumulate_grad_batches, cfg/kbr/default.yaml:122) and the AdamW kernels (bias correction and learning rate are host scalars).
    Data parallel (world > 1): with `opt` built on `gradient_buckets(model.nets)` each bucket's NCCL all-reduce is issued from an
    autograd hook the moment backward has finalised it and is captured INSIDE the graph on NCCL's stream, so the exchange overlaps
    with the rest of backward; if that capture is not possible the all-reduce runs after the replay (`overlap` tells which).

    `model` needs `.step(batch) -> (loss, ...)` and `.nets`; the reference's own MonoDepthModule qualifies after `plugin.install()`.
    The returned loss is a static tensor of the graph's pool: read (or copy) it before the next replay of ANY graph of the pool."""
    def __init__(self, model, opt, example_batch, warmup: int = 3, pool=None, accumulate: int = 1, overlap: bool | None = None):
        self.model, self.opt, self.accumulate, self._micro = model, opt, max(1, int(accumulate)), 0
        x, y, _ = example_batch
        clone = lambda d: {k: (v.clone() if torch.is_tensor(v) and v.is_cuda else v) for k, v in d.items()}
        self.static = (clone(x), clone(y), {})
        # Warm-up replays must not leave traces in the training state: BatchNorm running statistics (momentum updates) and the
        # tie-break noise counter are restored afterwards.
        buffers = [b for b in model.nets.buffers() if b.is_cuda]
        saved = [b.clone() for b in buffers]
        losses = getattr(model, 'losses', None)   # a dict here, an nn.ModuleDict in the reference's module
        crit = losses['img_recon'] if losses is not None and 'img_recon' in losses else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):   # lazy one-time initialisation must not happen under capture
                self.opt.zero_grad()
                self._fwd_bwd(False)
        torch.cuda.current_stream().wait_stream(side)
        step0 = crit.noise_step.clone() if crit is not None and getattr(crit, 'noise_step', None) is not None else None
        want_overlap = (opt.world > 1 and len(opt.buckets) > 1) if overlap is None else (overlap and opt.world > 1)
        self.overlap = False
        self.graph = None
        if want_overlap:
            try:
                self._capture(pool, True)
                self.overlap = True
            except Exception as e:   # e.g. an NCCL build that cannot be captured: fall back to the exchange after the replay
                self.overlap_error = f'{type(e).__name__}: {str(e)[:200]}'
                torch.cuda.synchronize()
                self.graph = None
        if self.graph is None: self._capture(pool, False)
        self.opt.zero_grad()
        for b, v in zip(buffers, saved): b.copy_(v)
        if step0 is not None: crit.noise_step.copy_(step0 - max(warmup, 1))

    def _capture(self, pool, overlap: bool) -> None:
        self.opt.zero_grad()
        graph = torch.cuda.CUDAGraph()
        # (thread_local: NCCL's watchdog thread may query events while the capture is open; the backward kernels are recorded
        # all the same — capture follows the stream, whichever thread launches into it)
        with torch.cuda.graph(graph, pool=pool, capture_error_mode='thread_local' if overlap else 'global'):
            self.loss = self._fwd_bwd(overlap)
        self.graph = graph

    def _fwd_bwd(self, overlap: bool) -> Tensor:
        self.model.bucket_hook = self.opt.bucket_ready if overlap else None
        try:
            loss = self.model.step(self.static)[0]
            loss.backward()
            if overlap:
                self.opt.all_reduce_async()    # the last (stem) bucket, and any whose hook did not fire
                self.opt.wait_all_reduce()     # joins NCCL's stream back into the capturing stream
        finally:
            self.model.bucket_hook = None
        return loss.detach()

    def load(self, batch) -> None:
        """Copy a batch (host-pinned or device tensors) into the graph's static input buffers, stream-ordered."""
        for dst, src in zip(self.static[:2], batch[:2]):
            for k, v in dst.items():
                if torch.is_tensor(v) and v.is_cuda: v.copy_(src[k], non_blocking=True)

    def replay(self, batch=None) -> Tensor:
        """Load + replay only (no gradient memset, exchange or optimiser step): the building block of `run` and of runners that
        own the accumulation themselves."""
        if batch is not None: self.load(batch)
        self.graph.replay()
        return self.loss

    def run(self, batch=None) -> Tensor:
        """One micro-batch. Every `accumulate`-th call ends with the gradient exchange (if not overlapped inside the graph) and the
        optimiser step on the mean of the accumulated gradients."""
        if self._micro == 0: self.opt.zero_grad()
        if batch is not None: self.load(batch)
        boundary = self._micro + 1 == self.accumulate
        if self.overlap and not boundary: raise RuntimeError('accumulate > 1 needs overlap=False (the captured exchange runs every replay).')
        self.graph.replay()
        self._micro += 1
        if boundary:
            if not self.overlap: self.opt.all_reduce_async()
            self.opt.step(grad_scale=1.0/self.accumulate)
            self._micro = 0
        return self.loss

    # -- pipelined input path: the next batch crosses PCIe on a side stream while the current step computes ---------------------
    def prefetch(self, batch) -> None:
        """Start the host->device copy of `batch` (pinned host tensors) into a staging buffer on the copy stream."""
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream()
            clone = lambda d: {k: (torch.empty_like(v) if torch.is_tensor(v) and v.is_cuda else v) for k, v in d.items()}
            self._staging = [(clone(self.static[0]), clone(self.static[1])) for _ in range(2)]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
            self._slot_w = self._slot_r = 0
        j = self._slot_w % 2
        self._copy_stream.wait_event(self._consumed[j])  # the step that read this staging slot has copied it out
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._staging[j], batch[:2]):
                for k, v in dst.items():
                    if torch.is_tensor(v) and v.is_cuda: v.copy_(src[k], non_blocking=True)
            self._ready[j].record(self._copy_stream)
        self._slot_w += 1

    def run_prefetched(self) -> Tensor:
        """Run one step on the oldest prefetched batch (device-to-device copy into the graph inputs, then replay)."""
        j = self._slot_r % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[j])
        for dst, src in zip(self.static[:2], self._staging[j]):
            for k, v in dst.items():
                if torch.is_tensor(v) and v.is_cuda: v.copy_(src[k], non_blocking=True)
        self._consumed[j].record(cur)
        self._slot_r += 1
        return self.run()


class ShapeCachedTrainStep:
    """Graph replay under the aspect-ratio augmentation (src/core/aspect_ratio.py; trainer.py:106), which changes the image
    size from step to step: one captured graph per distinct (b, n, H, W), built the first time a shape is seen and replayed
    afterwards. `sample_resize` only emits multiples of 32 with a bounded pixel count, so a run visits a few dozen shapes.
    The graphs never run concurrently and therefore share ONE private memory pool (the largest shape sets its size); libstv
    kernels are shape-agnostic (no per-shape compilation), so a new shape costs one eager warm-up + one capture.

    `max_graphs` bounds the cache (least recently used shape is dropped)."""
    def __init__(self, model: MonoDepthStep, opt, warmup: int = 2, max_graphs: int = 64, accumulate: int = 1):
        self.model, self.opt, self.warmup, self.max_graphs, self.accumulate = model, opt, warmup, max_graphs, accumulate
        self.steps: dict[tuple, GraphedTrainStep] = {}
        self.pool = None
        self._micro = 0

    @staticmethod
    def key(batch) -> tuple:
        x = batch[0]
        return (tuple(x['imgs'].shape), tuple(x['supp_imgs'].shape), tuple(int(i) for i in x['supp_idxs']))

    def run(self, batch) -> Tensor:
        k = self.key(batch)
        step = self.steps.pop(k, None)
        if step is None:
            if len(self.steps) >= self.max_graphs: self.steps.pop(next(iter(self.steps)))
            step = GraphedTrainStep(self.model, self.opt, batch, warmup=self.warmup, pool=self.pool, accumulate=1, overlap=False)
            if self.pool is None: self.pool = step.graph.pool()
        self.steps[k] = step  # most recently used last
        if self._micro == 0: self.opt.zero_grad()
        loss = step.replay(batch)
        self._micro += 1
        if self._micro == self.accumulate:   # micro-batches of different shapes accumulate into the same flat gradient buffer
            self.opt.all_reduce_async()
            self.opt.step(grad_scale=1.0/self.accumulate)
            self._micro = 0
        return loss
