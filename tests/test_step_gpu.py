"""GPU: one full training step of the product (networks + fused losses + backward, all libstv kernels) vs the oracle's
restatement of the reference step (oracle/step.py), shared weights, for the network / loss configurations BASELINE.json lists:
ResNet-18 / ResNet-18 (configs[1]), ConvNeXt-T + learned intrinsics + 4 support frames (configs[3]), ConvNeXt-B (configs[4]) —
at a size the oracle finishes in seconds.

Protocol. Every convolution / Linear of the product multiplies in TF32 on the tensor cores, which is the reference's own
numerics (`matmul: high`, cfg/default.yaml:171; cuDNN TF32). The yardstick is therefore measured, not assumed:
  ref  = the oracle in float64 on the CPU (exact arithmetic),
  lib  = the SAME oracle modules in float32 on the GPU with TF32 enabled (cuDNN / cuBLAS: what the reference would run),
  ours = the product.
All three get the same explicit tie-break noise and the product's per-pixel decisions (min-reprojection winner, auto-mask;
`forced_sel`), so the comparison is between smooth functions and a per-tensor bound is meaningful:
  loss:               |ours - ref| <= max(2 |lib - ref|, 1e-3 |ref|)   (cuDNN picks float32 kernels at this tiny size: TF32 floor)
  every parameter p:  ||g_ours - g_ref|| <= 2 ||g_lib - g_ref|| + 2e-3 ||g_ref||_global-scale floor (see `bound` below).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _grads(module):
    return {k: p.grad.detach().double().cpu() for k, p in module.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('depth_enc,n,learn_K', [('resnet18', 2, False), ('convnext_tiny', 4, True), ('convnext_base', 2, False)])
def test_training_step_matches_oracle(depth_enc, n, learn_K):
    import copy
    from oracle.step import OracleTrainer
    from slowtv_monodepth_b200 import synthetic as syn
    from slowtv_monodepth_b200.trainer import MonoDepthStep, default_cfg
    torch.manual_seed(11)
    ora = OracleTrainer(depth_enc, 'resnet18', learn_K=learn_K).double().train()
    with torch.no_grad():  # make the residual branches matter (layer-scale 1e-6 / zero-init BN would hide errors)
        for name, p in ora.nets.named_parameters():
            if name.endswith('gamma'): p.fill_(0.3)
            if name.endswith('bn2.weight'): p.fill_(1.0)
    model = MonoDepthStep(default_cfg(depth_enc, 'resnet18', learn_K=learn_K))
    model.nets.load_state_dict({k: v.float() for k, v in ora.nets.state_dict().items()})
    model = model.cuda().train()
    lib = copy.deepcopy(ora).float().cuda().train()
    b, shape, S = 2, (64, 96), 4
    batch = syn.make_batch(b, n, shape, seed=5)
    noise = torch.randn(S*b, 1, *shape, generator=torch.Generator().manual_seed(9))
    cast = lambda d, f: {k: (f(v) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
    b64 = (cast(batch[0], lambda v: v.double()), cast(batch[1], lambda v: v.double()), {})
    bgpu = (cast(batch[0], lambda v: v.cuda()), cast(batch[1], lambda v: v.cuda()), {})

    lp, _, _ = model.step(bgpu, noise=noise.cuda())
    lp.backward()
    torch.cuda.synchronize()
    sel = model.losses['img_recon'].last_sel.flatten(0, 1).unsqueeze(1)  # (S*b,1,H,W): the product's decisions

    lo, _, _ = ora.loss(b64, noise=noise.double(), forced_sel=sel.cpu())
    lo.backward()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True   # the reference's setting (trainer.py:30)
    try:
        ll, _, _ = lib.loss(bgpu, noise=noise.cuda(), forced_sel=sel)
        ll.backward()
        torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32

    go, gp, gl = _grads(ora.nets), _grads(model.nets), _grads(lib.nets)
    assert set(go) == set(gp) == set(gl)
    norm = lambda d: sum(float(v.pow(2).sum()) for v in d.values())**0.5
    diff = lambda a, c: {k: a[k] - c[k] for k in a}
    n_ref = norm(go)
    E_ours, E_lib = norm(diff(gp, go))/n_ref, norm(diff(gl, go))/n_ref
    e_loss, e_loss_lib = abs(lp.item() - lo.item())/abs(lo.item()), abs(ll.item() - lo.item())/abs(lo.item())
    print(f'{depth_enc} n={n} learn_K={learn_K}: loss {lp.item():.6f} (ref {lo.item():.6f}; rel err ours {e_loss:.2e}, lib {e_loss_lib:.2e}); '
          f'whole-gradient rel err ours {E_ours:.3e}, lib {E_lib:.3e}')
    assert torch.isfinite(lp) and e_loss <= max(2*e_loss_lib, 1e-3), (lp.item(), lo.item(), ll.item())
    assert E_ours <= max(2*E_lib, 2e-3), (E_ours, E_lib)
    # Per tensor: a tensor's error is bounded by twice the library's on the same tensor plus a floor of 2 x the library's
    # whole-gradient relative error applied to that tensor's own magnitude (tensors the library happens to get almost exactly
    # would otherwise set an unreachable bar), plus an absolute floor for gradients that are themselves rounding-sized.
    worst = []
    for k in go:
        ref_k = float(go[k].norm())
        e_o, e_l = float((gp[k] - go[k]).norm()), float((gl[k] - go[k]).norm())
        bound = 2*e_l + 2*max(E_lib, 1e-3)*ref_k + 1e-7*n_ref
        worst.append((e_o/bound, k, e_o, e_l, ref_k))
    worst.sort(reverse=True)
    for r, k, e_o, e_l, ref_k in worst[:5]: print(f'  {k}: err ours {e_o:.3e} lib {e_l:.3e} |g| {ref_k:.3e} -> {r:.2f} of the bound')
    assert worst[0][0] <= 1.0, worst[0]
