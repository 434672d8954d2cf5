"""CPU (gloo, world_size 2): the data-parallel logic of FlatAdamW — flat parameter/gradient buffers, one all-reduce of the
gradients, 1/world scaling folded into the optimiser — reproduces single-process training on the concatenated batch."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 4, 3, padding=1), nn.Flatten(), nn.Linear(4*8*8, 5))


def _data(rank=None):
    g = torch.Generator().manual_seed(123)
    x, y = torch.randn(8, 3, 8, 8, generator=g), torch.randn(8, 5, generator=g)
    return (x, y) if rank is None else (x[rank*4:(rank + 1)*4], y[rank*4:(rank + 1)*4])


def _train(model, opt, x, y, steps=3):
    for _ in range(steps):
        opt.zero_grad()
        ((model(x) - y)**2).mean().backward()
        opt.all_reduce_async()
        opt.step()
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def _train_bucketed(model, opt, x, y, steps=3, accumulate=1):
    """The overlapped protocol: autograd hooks start a bucket's all-reduce as soon as backward has finalised it (the node of the
    FIRST layer of the part that owns the bucket), the remaining bucket follows after backward; optional gradient accumulation."""
    for _ in range(steps):
        opt.zero_grad()
        for m in range(accumulate):
            h = model[2](model[1](model[0](x[m::accumulate])))
            h.grad_fn.register_hook(lambda *a: opt.bucket_ready(0) if m == accumulate - 1 else None)   # layers 2.. are final here
            ((model[4](model[3](h)) - y[m::accumulate])**2).mean().backward()
        opt.all_reduce_async()
        opt.step(grad_scale=1.0/accumulate)
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def _worker(rank, world, init_file, out_dir, bucketed=False):
    from slowtv_monodepth_b200.optim import FlatAdamW
    from tests import host_ref
    host_ref.install()  # spawned process: register the host reference arithmetic (the product path is CUDA-only)
    dist.init_process_group('gloo', init_method=f'file://{init_file}', rank=rank, world_size=world)
    try:
        model = _model()
        if rank == 1:   # a rank that starts from different weights is brought in line by the constructor's broadcast
            with torch.no_grad():
                for p in model.parameters(): p.add_(1.0)
        opt = FlatAdamW(model, lr=1e-2, weight_decay=1e-2, buckets=['4.', '2.'] if bucketed else None)
        assert opt.world == world and len(opt.buckets) == (3 if bucketed else 1)
        x, y = _data(rank)
        out = _train_bucketed(model, opt, x, y, accumulate=2) if bucketed else _train(model, opt, x, y)
        torch.save(out, os.path.join(out_dir, f'rank{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_overlap_and_accumulation_match_single_process():
    """Buckets in gradient-completion order, all-reduces started from autograd hooks, two accumulated micro-batches per step:
    same parameters as plain single-process training on the whole batch."""
    from slowtv_monodepth_b200.optim import FlatAdamW
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, 'init')
        mp.spawn(_worker, args=(2, init, d, True), nprocs=2, join=True)
        r0, r1 = torch.load(os.path.join(d, 'rank0.pt')), torch.load(os.path.join(d, 'rank1.pt'))
    assert torch.equal(r0, r1), 'ranks diverged'
    model = _model()
    ref = _train(model, FlatAdamW(model, lr=1e-2, weight_decay=1e-2), *_data())
    assert torch.allclose(r0, ref, atol=2e-6, rtol=1e-5), (r0 - ref).abs().max()


@pytest.mark.timeout(120)
def test_two_ranks_match_single_process():
    from slowtv_monodepth_b200.optim import FlatAdamW
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, 'init')
        mp.spawn(_worker, args=(2, init, d), nprocs=2, join=True)
        r0, r1 = torch.load(os.path.join(d, 'rank0.pt')), torch.load(os.path.join(d, 'rank1.pt'))
    assert torch.equal(r0, r1), 'ranks diverged'
    model = _model()
    opt = FlatAdamW(model, lr=1e-2, weight_decay=1e-2)
    ref = _train(model, opt, *_data())
    # mean over 8 samples == average of the two per-rank means over 4 samples each
    assert torch.allclose(r0, ref, atol=1e-6, rtol=1e-5), (r0 - ref).abs().max()


def test_flat_adamw_matches_torch_adamw():
    """Same update rule as torch.optim.AdamW with timm's no-decay rule for biases / 1-D parameters."""
    from slowtv_monodepth_b200.optim import FlatAdamW
    m1, m2 = _model(), _model()
    opt1 = FlatAdamW(m1, lr=1e-2, weight_decay=1e-2)
    decay = [p for n, p in m2.named_parameters() if p.ndim > 1]
    no_decay = [p for n, p in m2.named_parameters() if p.ndim <= 1]
    opt2 = torch.optim.AdamW([{'params': decay, 'weight_decay': 1e-2}, {'params': no_decay, 'weight_decay': 0.}], lr=1e-2)
    x, y = _data()
    for _ in range(4):
        opt1.zero_grad(); ((m1(x) - y)**2).mean().backward(); opt1.step()
        opt2.zero_grad(); ((m2(x) - y)**2).mean().backward(); opt2.step()
    for (n, a), b in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5), n


def test_lr_schedule_matches_the_reference_chained_schedulers():
    """`LRSchedule` (closed form) vs torch's ChainedScheduler([StepLR, LinearLR]) stepped per epoch, as the reference builds it
    (src/core/trainer.py:85-94 with cfg/kbr/default.yaml's scheduler block)."""
    import pytest
    import torch
    from torch.optim.lr_scheduler import ChainedScheduler, LinearLR, StepLR
    from slowtv_monodepth_b200.optim import LRSchedule
    for cfg in ({'steplr': {'step_size': 40, 'gamma': 0.1}, 'linear': {'start_factor': 0.1, 'total_iters': 4}},
                {'steplr': {'step_size': 3, 'gamma': 0.5}}, {'linear': {'start_factor': 0.25, 'end_factor': 0.75, 'total_iters': 6}}, {}):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=1e-4)
        scheds = []
        if 'steplr' in cfg: scheds.append(StepLR(opt, **cfg['steplr']))
        if 'linear' in cfg: scheds.append(LinearLR(opt, **cfg['linear']))
        chained = ChainedScheduler(scheds) if scheds else None
        ours = LRSchedule(1e-4, cfg)
        for epoch in range(100):
            assert ours.lr(epoch) == pytest.approx(opt.param_groups[0]['lr'], rel=1e-9), (cfg, epoch)
            opt.step()
            if chained: chained.step()
    with pytest.raises(KeyError): LRSchedule(1e-4, {'cosine': {}})
