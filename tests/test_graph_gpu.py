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
    g.load(b1)
    g.graph.replay()
    torch.cuda.synchronize()
    loss_g, grad_g = g.loss.clone(), opt.grad.clone()
    opt.zero_grad()
    model.losses['img_recon']._calls -= 1  # same tie-break noise seed as the captured call
    loss_e, _, _ = model.step(b1)
    loss_e.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss_g) and abs(loss_g.item() - loss_e.item()) <= 1e-6*abs(loss_e.item()) + 1e-7
    err = (grad_g - opt.grad).norm()/opt.grad.norm()
    assert err < 1e-5, err.item()   # split-K / atomics reorder fp32 sums; everything else is bit-identical
