"""GPU: network kernels (through the C ABI) and the networks built on them vs the oracle restatements (float64 on CPU).
Convolutions / GEMMs run in TF32 on the tensor cores, like the reference (`matmul: high`, cfg/default.yaml:171), so whole-
network comparisons use TF32-appropriate bounds; the fp32 kernels (depthwise conv, LayerNorm) are held to 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(2, 24, 40, 96), (1, 7, 9, 33), (2, 12, 20, 192), (1, 6, 10, 160), (1, 45, 21, 40), (3, 3, 3, 8)])
def test_dwconv7_matches_oracle(shape):
    from slowtv_monodepth_b200 import functional as F_
    N, H, W, C = shape
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, H, W, C, generator=g, dtype=torch.float64)
    w = torch.randn(C, 1, 7, 7, generator=g, dtype=torch.float64)*0.2
    b = torch.randn(C, generator=g, dtype=torch.float64)
    gy = torch.randn(N, H, W, C, generator=g, dtype=torch.float64)

    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    yr = F.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=3, groups=C).permute(0, 2, 3, 1)
    yr.backward(gy)
    xc, wc, bc = (t.float().cuda().requires_grad_() for t in (x, w, b))
    yc = F_.dwconv7(xc, wc, bc)
    yc.backward(gy.float().cuda())
    assert U.rel(yc, yr) < 1e-5
    assert U.rel(xc.grad, xr.grad) < 1e-5 and U.rel(wc.grad, wr.grad) < 1e-5 and U.rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize('P,C', [(100, 96), (37, 33), (64, 768), (50, 1024), (9, 160), (33, 192), (21, 384), (17, 512), (40, 4)])
def test_layernorm_matches_oracle(P, C):
    from slowtv_monodepth_b200 import functional as F_
    g = torch.Generator().manual_seed(1)
    x = torch.randn(P, C, generator=g, dtype=torch.float64)*2 + 0.5
    ga, be = torch.randn(C, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    gy = torch.randn(P, C, generator=g, dtype=torch.float64)
    xr, gr, br = (t.clone().requires_grad_() for t in (x, ga, be))
    F.layer_norm(xr, (C,), gr, br, 1e-6).backward(gy)
    xc, gc, bc = (t.float().cuda().requires_grad_() for t in (x, ga, be))
    yc = F_.layer_norm(xc, gc, bc, 1e-6)
    yc.backward(gy.float().cuda())
    assert U.rel(yc, F.layer_norm(x, (C,), ga, be, 1e-6)) < 1e-5
    assert U.rel(xc.grad, xr.grad) < 2e-5 and U.rel(gc.grad, gr.grad) < 1e-5 and U.rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize('enc', ['resnet18', 'convnext_tiny'])
def test_networks_match_oracle_networks(enc):
    """Product DepthNet / PoseNet on the GPU vs the oracle's restatement in float64 on the CPU, shared weights.
    Every convolution / Linear layer of the product runs in TF32 on the tensor cores (the reference's own setting, `matmul:
    high`), so the bounds are TF32 bounds: ~1e-3 per layer output, growing through the ~60-layer backward. The kernels'
    indexing logic itself is pinned bit-exactly by tests/test_conv_gpu.py and tests/test_gemm_gpu.py."""
    from oracle import nets as ON
    from slowtv_monodepth_b200.networks import DepthNet, PoseNet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(2)
        od, op = ON.DepthNet(enc).double(), ON.PoseNet('resnet18', learn_K=True).double()
        with torch.no_grad():  # make the residual branches matter (layer-scale 1e-6 / zero-init BN would hide errors)
            for n, p in od.named_parameters():
                if n.endswith('gamma'): p.fill_(0.5)
                if n.endswith('bn2.weight'): p.fill_(1.0)
        pd, pp = DepthNet(enc, pretrained=False), PoseNet('resnet18', learn_K=True)
        pd.load_state_dict({k: v.float() for k, v in od.state_dict().items()})
        pp.load_state_dict({k: v.float() for k, v in op.state_dict().items()})
        pd, pp = pd.cuda().train(), pp.cuda().train()
        x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
        a, b = od(x), pd(x.float().cuda())
        for s in range(4): assert U.rel(b['disp'][s], a['disp'][s]) < 3e-3, (s, U.rel(b['disp'][s], a['disp'][s]))
        la = sum((v*v).sum() for v in a['disp'].values()); lb = sum((v*v).sum() for v in b['disp'].values())
        ga = dict(zip([n for n, _ in od.named_parameters()], torch.autograd.grad(la, list(od.parameters()))))
        gb = dict(zip([n for n, _ in pd.named_parameters()], torch.autograd.grad(lb, list(pd.parameters()))))
        # TF32 rounding (~1e-3 per layer output) does not cancel in long sums, so a gradient that is itself a small difference
        # of large terms (BatchNorm biases) carries a large RELATIVE error in any TF32 implementation (cuDNN's included):
        # the bound is on the error relative to the whole gradient, plus a loose per-tensor cap that catches wiring mistakes.
        num = sum(float((gb[n].double().cpu() - ga[n]).pow(2).sum()) for n in ga)**0.5
        den = sum(float(ga[n].pow(2).sum()) for n in ga)**0.5
        assert num/den < 1e-2, num/den
        worst = max((U.rel(gb[n], ga[n]), n) for n in ga if ga[n].abs().max() > 1e-12)
        assert worst[0] < 0.3, worst
        x6 = torch.randn(2, 6, 64, 96, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
        a, b = op(x6), pp(x6.float().cuda())
        for k in ('R', 't', 'fs', 'cs'): assert U.rel(b[k], a[k]) < 3e-3, (k, U.rel(b[k], a[k]))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True


@pytest.mark.parametrize('shape,relu,with_res', [((2, 6, 10, 64), True, True), ((3, 5, 7, 128), True, False), ((2, 4, 4, 512), False, True),
                                                 ((4, 9, 11, 32), False, False)])
def test_batchnorm_matches_torch(shape, relu, with_res):
    """stv_bn_fwd / stv_bn_bwd (train mode, fused residual + ReLU) vs nn.functional.batch_norm in float64."""
    from slowtv_monodepth_b200 import functional as F_
    g = torch.Generator().manual_seed(7)
    C = shape[-1]
    x = torch.randn(shape, generator=g, dtype=torch.float64)*1.5 + 0.3
    res = torch.randn(shape, generator=g, dtype=torch.float64) if with_res else None
    ga, be = torch.rand(C, generator=g, dtype=torch.float64) + 0.5, torch.randn(C, generator=g, dtype=torch.float64)
    gy = torch.randn(shape, generator=g, dtype=torch.float64)
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)

    leaves = [t.clone().requires_grad_() for t in (x, ga, be)] + ([res.clone().requires_grad_()] if with_res else [])
    yr = F.batch_norm(leaves[0].permute(0, 3, 1, 2), rm, rv, leaves[1], leaves[2], True, 0.1, 1e-5).permute(0, 2, 3, 1)
    if with_res: yr = yr + leaves[3]
    if relu: yr = torch.relu(yr)
    yr.backward(gy)

    cl = [t.float().cuda().requires_grad_() for t in (x, ga, be)] + ([res.float().cuda().requires_grad_()] if with_res else [])
    rmc, rvc = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
    yc = F_.batch_norm_nhwc(cl[0], cl[1], cl[2], res=cl[3] if with_res else None, relu=relu, run_mean=rmc, run_var=rvc)
    yc.backward(gy.float().cuda())
    assert U.rel(yc, yr) < 1e-5
    assert U.rel(rmc, rm) < 1e-5 and U.rel(rvc, rv) < 1e-5
    for a, b in zip(cl, leaves): assert U.rel(a.grad, b.grad) < 2e-5


@pytest.mark.parametrize('shape', [(2, 12, 20, 64), (1, 7, 9, 8), (3, 6, 6, 32)])
def test_maxpool_matches_torch(shape):
    """stv_maxpool3x3s2_fwd/bwd vs F.max_pool2d(3, 2, 1) (values, and gradients given distinct inputs)."""
    from slowtv_monodepth_b200 import functional as F_
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)
    gy_shape = (shape[0], (shape[1] - 1)//2 + 1, (shape[2] - 1)//2 + 1, shape[3])
    gy = torch.randn(gy_shape, generator=g)
    xr = x.clone().requires_grad_()
    yr = F.max_pool2d(xr.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    yr.backward(gy)
    xc = x.cuda().requires_grad_()
    yc = F_.maxpool3x3s2(xc)
    yc.backward(gy.cuda())
    assert torch.equal(yc.cpu(), yr.detach()) and torch.allclose(xc.grad.cpu(), xr.grad, atol=1e-6)


@pytest.mark.parametrize('shapes', [[(8, 1, 384, 640)]*8, [(3, 1, 37, 53), (1, 5), (2, 3, 4, 5, 6)], [(1, 1, 33, 65)]*17])
def test_mean_std_matches_torch(shapes):
    """stv_mean_std (the logging statistics of trainer.py:486-503) vs torch.mean / torch.std (unbiased), incl. a map far from zero."""
    from slowtv_monodepth_b200 import functional as F_
    torch.manual_seed(0)
    ts = [torch.rand(s, device='cuda')*(1 + 30*(i % 3)) + 40*(i % 2) for i, s in enumerate(shapes)]
    got = F_.mean_std(ts).cpu().double()
    want = torch.tensor([[t.double().mean().item(), t.double().std().item()] for t in ts], dtype=torch.float64)
    assert torch.allclose(got, want, rtol=2e-6, atol=1e-7), (got - want).abs().max()


def test_step_summary_has_the_reference_keys_and_values():
    from slowtv_monodepth_b200 import synthetic as syn
    from slowtv_monodepth_b200.trainer import MonoDepthStep, default_cfg, summarize
    torch.manual_seed(0)
    model = MonoDepthStep(default_cfg('resnet18', 'resnet18', learn_K=True)).cuda().train()
    batch = syn.make_batch(2, 2, (64, 96), seed=0, device='cuda')
    with torch.no_grad(): _, _, fwd = model.step(batch, want_up=True)
    got = summarize(fwd).to_host()
    for s in range(4):
        for key in ('disp', 'depth'):
            v = fwd[f'{key}_up'][s]
            assert abs(got[f'{key}_mean_{s}'] - v.mean().item()) <= 1e-5*abs(v.mean().item()) + 1e-7
            assert abs(got[f'{key}_std_{s}'] - v.std().item()) <= 1e-4*abs(v.std().item()) + 1e-7
    assert {'T_-1_t_mean', 'T_1_R_std', 'fx', 'cy'} <= set(got)
