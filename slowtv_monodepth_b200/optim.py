"""Flat-buffer AdamW + gradient all-reduce for the data-parallel training step.

Reference behaviour being replaced: `timm.optim.create_optimizer_v2(nets, opt='adamw', lr, weight_decay)`
(src/tools/parsers.py:205-243) -> torch.optim.AdamW(foreach) with timm's rule that biases and 1-D parameters get no weight
decay, and Lightning's DDP gradient averaging (api/train/train.py:105-106).

Here every parameter of every network lives in ONE contiguous fp32 buffer (decayed parameters first), gradients in a
second one, so that (a) the optimiser is a single libstv kernel launch and (b) the data-parallel exchange is a single NCCL
all-reduce over NVLink on a side stream, with no per-tensor bookkeeping on the host.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib as L_
from . import functional as F_

__all__ = ['FlatAdamW', 'LRSchedule']


class LRSchedule:
    """Per-epoch learning rate of the reference's chained schedulers, in closed form.

    Reference: `configure_optimizers` (src/core/trainer.py:85-94) wraps every entry of `cfg['scheduler']` (src/tools/parsers.py:246-269;
    KBR: `steplr` {step_size 40, gamma 0.1} + `linear` {start_factor 0.1, total_iters 4}, cfg/kbr/default.yaml:97-103) in one
    `torch.optim.lr_scheduler.ChainedScheduler`, stepped once per epoch by Lightning. The chained recursive updates multiply out
    to  lr(e) = base * gamma^floor(e/step_size) * (start + (end - start) * min(e, total_iters)/total_iters).
    FlatAdamW's learning rate is a host scalar handed to the kernel, so `apply(opt, epoch)` is all a training loop needs."""
    def __init__(self, base_lr: float, cfg: dict | None):
        self.base_lr, self.factors = float(base_lr), []
        for name, kw in (cfg or {}).items():
            if kw is None: continue
            if name == 'steplr':
                step, gamma = int(kw['step_size']), float(kw.get('gamma', 0.1))
                if step <= 0: raise ValueError(f'steplr: step_size must be positive (got {step})')
                self.factors.append(lambda e, step=step, gamma=gamma: gamma**(e//step))
            elif name == 'linear':
                start, end, total = float(kw.get('start_factor', 1/3)), float(kw.get('end_factor', 1.0)), int(kw.get('total_iters', 5))
                if not 0 < start <= 1 or not 0 <= end <= 1: raise ValueError('linear: factors must lie in (0, 1] / [0, 1]')
                self.factors.append(lambda e, start=start, end=end, total=total: start + (end - start)*min(e, total)/total)
            else:
                raise KeyError(f'Unsupported scheduler "{name}" (steplr | linear).')

    def lr(self, epoch: int) -> float:
        out = self.base_lr
        for f in self.factors: out *= f(int(epoch))
        return out

    def apply(self, opt: 'FlatAdamW', epoch: int) -> float:
        opt.lr = self.lr(epoch)
        return opt.lr


class FlatAdamW:
    """Every trainable parameter in one flat fp32 buffer + fused AdamW + one gradient all-reduce.

    Differences from torch.optim.AdamW worth knowing: a parameter that took no part in the step has a ZERO gradient here (the
    buffer is zeroed, never set to None), so it still receives weight decay, where torch skips `grad is None` parameters; on
    the hot path every parameter of both networks is used every step (the unused half of the pose head's output channels gets an
    exact-zero, not-None gradient in the reference too), so the two agree — `tests/test_optim_gpu.py` holds the kernel to
    torch.optim.AdamW with timm's no-decay groups.
    With world > 1 the constructor broadcasts rank 0's parameters (DDP's initial sync); `broadcast_buffers()` does the same for
    module buffers (BatchNorm running statistics) and should be called after loading a checkpoint on one rank."""
    def __init__(self, module: nn.Module, lr: float = 1e-4, weight_decay: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 channels_last: bool = True, buckets: list[str] | None = None):
        """buckets: parameter-name prefixes in the order their gradients COMPLETE during backward (e.g. ['pose.',
        'depth.decoders.', 'depth.encoder.stages_3', ...]); parameters matching no prefix form a last bucket. Each bucket is one
        contiguous slice of the flat buffers (its decayed parameters first), so a bucket is one all-reduce + one AdamW launch and
        `bucket_ready(j)` can start bucket j's all-reduce while backward is still producing the later ones."""
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named: raise ValueError('No trainable parameters.')
        no_decay = lambda n, p: p.ndim <= 1 or n.endswith('.bias')  # timm `param_groups_weight_decay`
        prefixes = list(buckets or [])
        which = lambda n: next((j for j, pre in enumerate(prefixes) if n.startswith(pre)), len(prefixes))
        groups: list[list] = [[] for _ in range(len(prefixes) + 1)]
        for n, p in named: groups[which(n)].append((n, p))
        groups = [g for g in groups if g]
        dev = named[0][1].device
        # Every parameter starts on a 16-byte boundary (TMA / 128-bit loads read weights and biases in place); padding stays zero.
        al = lambda n: (n + 3)//4*4
        self.params, self.buckets = [], []   # buckets: (start, size, n_decay) in elements of the flat buffers
        total = 0
        for g in groups:
            dec = [p for n, p in g if not no_decay(n, p)]
            nod = [p for n, p in g if no_decay(n, p)]
            n_dec, n_all = sum(al(p.numel()) for p in dec), sum(al(p.numel()) for p in dec + nod)
            self.buckets.append((total, n_all, n_dec))
            self.params += dec + nod
            total += n_all
        self.n_decay = self.buckets[0][2] if len(self.buckets) == 1 else None
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.ndim == 4 and channels_last:
                # Convolution filters live in the flat buffer in (O, kh, kw, I) order, exposed as channels-last (O, I, kh, kw)
                # tensors: the NHWC implicit-GEMM kernels then read filters and write filter gradients in place, with no
                # per-step layout conversion (617 nhwc<->nchw transposes, ~10 ms, per step otherwise).
                o, i, kh, kw = p.shape
                view = lambda buf: buf[off:off + n].view(o, kh, kw, i).permute(0, 3, 1, 2)
            else:
                view = lambda buf: buf[off:off + n].view(p.shape)
            view(self.flat).copy_(p.data)
            p.data = view(self.flat)
            p.grad = view(self.grad)
            off += al(n)
        if self.flat.is_cuda: F_.grad_sink(True)  # gradients live in one pre-zeroed flat buffer: kernels accumulate into it in place
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._works: list = []
        self._reduced: set[int] = set()
        self._module = module
        if self.world > 1: dist.broadcast(self.flat, 0)

    def broadcast_buffers(self, src: int = 0) -> None:
        """Rank `src`'s parameters and module buffers to every rank (after a checkpoint load on one rank)."""
        if self.world <= 1: return
        dist.broadcast(self.flat, src)
        for buf in self._module.buffers():
            if buf.numel(): dist.broadcast(buf, src)

    def zero_grad(self) -> None:
        """Gradients are accumulated in place into the flat buffer, so they are zeroed (one memset), not set to None."""
        self.grad.zero_()

    def bucket_ready(self, j: int) -> None:
        """Start the sum-all-reduce of bucket j (its gradients are final). Called from autograd hooks while backward is still
        running the networks whose gradients complete later; the collective runs on NCCL's own stream, ordered after the kernels
        enqueued so far, so it overlaps with the rest of backward (and is captured with it when the step is a CUDA graph)."""
        if self.world <= 1 or j in self._reduced or j >= len(self.buckets): return
        start, n, _ = self.buckets[j]
        self._reduced.add(j)
        if self.grad.is_cuda: F_.join_side_stream(self.grad.device)   # weight gradients enqueued on the side stream so far
        self._works.append(dist.all_reduce(self.grad[start:start + n], op=dist.ReduceOp.SUM, async_op=True))

    def all_reduce_async(self):
        """Sum-all-reduce of every bucket not started yet (the 1/world scale is folded into the optimiser kernel)."""
        for j in range(len(self.buckets)): self.bucket_ready(j)

    def wait_all_reduce(self) -> None:
        """Make the current stream wait for the outstanding all-reduces (no host synchronisation with NCCL)."""
        for w in self._works: w.wait()
        self._works.clear()
        self._reduced.clear()

    def step(self, grad_scale: float = 1.0) -> None:
        """grad_scale: extra factor on the gradient (1/k after accumulating k micro-batches, as Lightning divides the loss)."""
        self.wait_all_reduce()
        self.step_count += 1
        if self.flat.is_cuda:
            for start, n, n_dec in self.buckets:
                sl = slice(start, start + n)
                F_.adamw_step_(self.flat[sl], self.grad[sl], self.exp_avg[sl], self.exp_avg_sq[sl], n_decay=n_dec, lr=self.lr,
                               beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay,
                               step=self.step_count, grad_scale=grad_scale/self.world)
        else:
            L_.host_path(self, grad_scale, what='FlatAdamW.step')
