"""GPU: the CUDA-graph training step replays exactly what the eager step computes (same loss, same gradients)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_step_matches_eager():
    from slowtv_monodepth_b200 import synthetic as syn
    from slowtv_monodepth_b200.optim import FlatAdamW
    from slowtv_monodepth_b200.trainer import GraphedTrainStep, MonoDepthStep, default_cfg
    torch.manual_seed(0)
    model = MonoDepthStep(default_cfg('convnext_tiny', 'resnet18', learn_K=True)).cuda().train()
    opt = FlatAdamW(model.nets)
    b0 = syn.make_batch(2, 2, (64, 96), seed=0, device='cuda')
    b1 = syn.make_batch(2, 2, (64, 96), seed=1, device='cuda')
    g = GraphedTrainStep(model, opt, b0)
    # replay on a NEW batch, then the same batch eagerly: the flat gradient buffers must agree
    crit = model.losses['img_recon']
    g.load(b1)
    opt.zero_grad()             # the gradient memset lives outside the graph (micro-batches may accumulate)
    crit.noise_step.fill_(41)   # the tie-break noise is seeded by a DEVICE counter that every call (and replay) advances
    g.graph.replay()
    torch.cuda.synchronize()
    assert crit.noise_step.item() == 42   # the replay drew its noise and advanced the counter, like randn_like would
    loss_g, grad_g = g.loss.clone(), opt.grad.clone()
    opt.zero_grad()
    crit.noise_step.fill_(41)   # same draw for the eager step
    loss_e, _, _ = model.step(b1)
    loss_e.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss_g) and abs(loss_g.item() - loss_e.item()) <= 1e-6*abs(loss_e.item()) + 1e-7
    err = (grad_g - opt.grad).norm()/opt.grad.norm()
    assert err < 1e-5, err.item()   # split-K / atomics reorder fp32 sums; everything else is bit-identical


def test_shape_cached_graphs_follow_the_aspect_ratio_augmentation():
    """Augmented batches change size every step: each distinct shape gets its own captured graph (shared memory pool); replaying
    a cached shape reproduces the eager step on the same batch."""
    import random
    from slowtv_monodepth_b200 import aspect_ratio as AR, synthetic as syn
    from slowtv_monodepth_b200.optim import FlatAdamW
    from slowtv_monodepth_b200.trainer import MonoDepthStep, ShapeCachedTrainStep, default_cfg
    torch.manual_seed(0)
    model = MonoDepthStep(default_cfg('convnext_tiny', 'resnet18')).cuda().train()
    opt = FlatAdamW(model.nets)
    runner = ShapeCachedTrainStep(model, opt, warmup=1)
    shapes, batches = [], []
    for seed in range(4):
        random.seed(seed); torch.manual_seed(seed)
        b = AR.aspect_ratio_aug(syn.make_batch(2, 2, (128, 192), seed=seed, device='cuda'), p=1.0, ref_shape=(128, 192))
        assert all(s % 32 == 0 for s in b[0]['imgs'].shape[-2:]) and b[0]['imgs'].shape[-2:] == b[1]['supp_imgs'].shape[-2:]
        shapes.append(tuple(b[0]['imgs'].shape[-2:])); batches.append(b)
    for b in batches + batches[:2]:
        loss = runner.run(b)
        assert torch.isfinite(loss)
    assert len(runner.steps) == len(set(shapes)) and len(set(shapes)) >= 2
    # replay of a cached shape (graph only, no optimiser step) == eager step on the same batch with the same parameters
    b = batches[0]
    step = runner.steps[runner.key(b)]
    crit = model.losses['img_recon']
    crit.noise_step.fill_(7)
    opt.zero_grad()   # the gradient memset lives outside the graph (micro-batches may accumulate)
    step.load(b); step.graph.replay(); torch.cuda.synchronize()
    loss_g, grad_g = step.loss.clone(), opt.grad.clone()
    opt.zero_grad()
    crit.noise_step.fill_(7)    # same tie-break draw as the replay (the counter lives on the device, shared by every graph)
    loss_e, _, _ = model.step(b)
    loss_e.backward()
    torch.cuda.synchronize()
    assert abs(loss_g.item() - loss_e.item()) <= 1e-6*abs(loss_e.item()) + 1e-7
    assert ((grad_g - opt.grad).norm()/opt.grad.norm()).item() < 1e-5
