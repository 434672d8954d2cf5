"""GPU: the drop-in boundary exercised on the device through the REFERENCE'S OWN module.

The reference travels to the GPU box as the byte-compiled build `oracle/_ref/` (oracle/build_ref.py; unmodified, compiled from
the sources where they lie). Its `MonoDepthModule(cfg)` is built through its own parsers / registry and its own `step`,
`forward_loss` and handlers run on CUDA:
  (a) loss side only   install(nets=False): identical (reference, cuDNN) networks on both sides, so the two runs see bit-identical
      disparities and poses; the loss must agree to 1e-5 and the parameter gradients to the float32 sub-gradient noise floor
      the reference shows against itself (tests/test_oracle_golden.py::test_fp32_reference_noise_floor);
  (b) everything       install(): B200 networks + loss + sync-free step; TF32 bounds (both sides multiply in TF32);
  (c) the reference's step captured as one CUDA graph replays what the eager call computes;
  (d) `ReconstructionLoss.forward` on pre-warped frames (the registered class called directly) against the reference's class.
"""
import copy
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.available(), reason='reference build (oracle/_ref) not present')]

CFG = {
    'net': {'depth': {'enc_name': 'convnext_tiny', 'pretrained': False, 'dec_name': 'monodepth', 'out_scales': [0, 1, 2, 3]},
            'pose': {'enc_name': 'resnet18', 'pretrained': False, 'learn_K': True}},
    'loss': {'img_recon': {'weight': 1, 'loss_name': 'ssim', 'use_min': True, 'use_automask': True},
             'disp_smooth': {'weight': 0.001, 'use_edges': True}},
    'optimizer': {'type': 'adamw', 'lr': 1e-4}, 'scheduler': None, 'dataset': {}, 'loader': {'batch_size': 2},
    'trainer': {'min_depth': 0.1, 'max_depth': 100, 'always_fwd_pose': True},
}


def _modules(install_kwargs, batch=None, exact_geometry=False):
    """-> (reference module built from the reference's classes, reference module built after install(), plugin[, the pure
    reference's step result]) with shared weights. `install()` rebinds class-level names, so the pure reference step (when a
    batch is given) runs BEFORE it."""
    warnings.filterwarnings('ignore')
    ref_shim.load()
    import src.core.trainer as rt
    from slowtv_monodepth_b200 import plugin
    torch.manual_seed(3)
    plugin.uninstall()
    ref = rt.MonoDepthModule(copy.deepcopy(CFG))
    with torch.no_grad():
        for name, p in ref.nets.named_parameters():
            if name.endswith('gamma'): p.fill_(0.3)
            if name.endswith('bn2.weight'): p.fill_(1.0)
    ref = ref.cuda().train()
    ref_out = None
    if batch is not None:
        # MonoDepthModule.__init__ sets float32 matmuls to TF32 (trainer.py:30), which on a GPU also rounds the reference's own
        # geometry (the K^-1 / T / K `bmm`s of ViewSynth, geometry.py:313,386,341) to 10 mantissa bits. `exact_geometry` restores
        # float32 matmuls for the comparison that isolates the loss side.
        if exact_geometry: torch.set_float32_matmul_precision('highest')
        torch.backends.cudnn.deterministic = True
        try: ref_out = _step(ref, batch)
        finally: torch.backends.cudnn.deterministic = False
    plugin.install(**install_kwargs)
    ours = rt.MonoDepthModule(copy.deepcopy(CFG))
    ours.nets.load_state_dict(ref.nets.state_dict())
    return ref, ours.cuda().train(), plugin, ref_out


def _step(module, batch):
    for p in module.nets.parameters(): p.grad = None
    loss, ld, fwd = module.step(batch, mode='train')
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach(), ld, {k: p.grad.detach().double() for k, p in module.nets.named_parameters() if p.grad is not None}


def _gerr(a, b):
    num = sum(float((a[k] - b[k]).pow(2).sum()) for k in b)**0.5
    return num/sum(float(b[k].pow(2).sum()) for k in b)**0.5


@pytest.fixture
def batch():
    from slowtv_monodepth_b200 import synthetic as syn
    return syn.make_batch(2, 2, (64, 96), seed=5, device='cuda')


def test_loss_side_drop_in_matches_the_reference_step(batch):
    ref, ours, plugin, (l_ref, ld_ref, g_ref) = _modules(dict(nets=False, loss=True), batch, exact_geometry=True)
    try:
        from slowtv_monodepth_b200 import losses
        assert isinstance(ours.losses['img_recon'], losses.ReconstructionLoss) and type(ours.nets['depth']) is type(ref.nets['depth'])
        torch.backends.cudnn.deterministic = True
        l_our, ld_our, g_our = _step(ours, batch)
        assert abs(l_our.item() - l_ref.item()) <= 1e-5*abs(l_ref.item()), (l_our.item(), l_ref.item())
        for k in ('loss_img_recon', 'loss_disp_smooth'):
            assert abs(ld_our[k].item() - ld_ref[k].item()) <= 1e-5*abs(ld_ref[k].item()), k
        assert (ld_our['automask'] != ld_ref['automask']).float().mean().item() < 5e-3
        assert torch.allclose(ld_our['supp_imgs_warp'], ld_ref['supp_imgs_warp'], atol=2e-5)
        assert set(g_our) == set(g_ref)
        e = _gerr(g_our, g_ref)
        print(f'loss-side drop-in: loss {l_our.item():.7f} vs {l_ref.item():.7f}, whole-gradient rel diff {e:.3e}')
        assert e < 3e-2, e   # float32 sub-gradient events (L1 sign / texel cell) differ between ANY two float32 evaluations
    finally:
        torch.backends.cudnn.deterministic = False
        torch.set_float32_matmul_precision('high')
        plugin.uninstall()


def test_full_drop_in_matches_the_reference_step_within_tf32(batch):
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True   # the reference's numerics (trainer.py:30)
    ref, ours, plugin, (l_ref, _, g_ref) = _modules({}, batch)
    try:
        from slowtv_monodepth_b200 import networks
        assert isinstance(ours.nets['depth'], networks.DepthNet) and isinstance(ours.nets['pose'], networks.PoseNet)
        l_our, ld, g_our = _step(ours, batch)
        assert set(g_our) == set(g_ref)
        e = _gerr(g_our, g_ref)
        print(f'full drop-in: loss {l_our.item():.6f} vs {l_ref.item():.6f}, whole-gradient rel diff {e:.3e}')
        assert abs(l_our.item() - l_ref.item()) <= 2e-3*abs(l_ref.item())
        assert e < 0.15, e   # two TF32 evaluations + decision flips at near-ties; the forced-decision bound is tests/test_step_gpu.py
        assert {'supp_imgs_warp', 'automask', 'disp_grad', 'image_grad'} <= set(ld)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        plugin.uninstall()


def test_reference_step_replays_as_a_cuda_graph(batch):
    from slowtv_monodepth_b200.optim import FlatAdamW
    ref, ours, plugin, _ = _modules({})
    try:
        opt = FlatAdamW(ours.nets)
        runner = plugin.graphed_step(ours, opt, batch, warmup=2)
        crit = ours.losses['img_recon']
        crit.noise_step.fill_(5)
        runner.load(batch); runner.graph.replay(); torch.cuda.synchronize()
        loss_g, grad_g = runner.loss.clone(), opt.grad.clone()
        opt.zero_grad()
        crit.noise_step.fill_(5)
        loss_e, _, _ = ours.step(batch, mode='train')
        loss_e.backward()
        torch.cuda.synchronize()
        assert abs(loss_g.item() - loss_e.item()) <= 1e-6*abs(loss_e.item()) + 1e-7
        assert ((grad_g - opt.grad).norm()/opt.grad.norm()).item() < 1e-5
        before = opt.flat.clone()
        assert torch.isfinite(runner.run(batch)) and (opt.flat != before).any()
    finally:
        plugin.uninstall()


@pytest.mark.parametrize('use_min,use_automask,loss_name', [(True, True, 'ssim'), (False, False, 'ssim'), (True, False, 'l1'), (False, True, 'ssim')])
def test_reconstruction_loss_forward_on_prewarped_frames(use_min, use_automask, loss_name):
    """The registered `img_recon` class called directly (reconstruction.py:98-126) vs the reference's class, float64 on the CPU."""
    ref_shim.load()
    from src import losses as ref_losses
    from slowtv_monodepth_b200.losses import ReconstructionLoss
    from slowtv_monodepth_b200 import synthetic as syn
    imgs, supp = syn.make_frames(2, 3, (40, 56), seed=8)
    g = torch.Generator().manual_seed(1)
    pred = (supp + 0.03*torch.randn(supp.shape, generator=g)).clamp(0, 1)
    noise = torch.randn(2, 1, 40, 56, generator=g)
    rc = ref_losses.ReconstructionLoss(loss_name=loss_name, use_min=use_min, use_automask=use_automask)
    p64 = pred.double().requires_grad_()
    orig = torch.randn_like
    torch.randn_like = lambda x: noise.double()
    try: l_ref, ld_ref = rc(p64, imgs.double(), source=supp.double())
    finally: torch.randn_like = orig
    l_ref.backward()
    oc = ReconstructionLoss(loss_name=loss_name, use_min=use_min, use_automask=use_automask)
    pc = pred.cuda().requires_grad_()
    l_our, ld_our = oc(pc, imgs.cuda(), source=supp.cuda(), noise=noise.cuda())
    l_our.backward()
    torch.cuda.synchronize()
    assert abs(l_our.item() - l_ref.item()) <= 1e-5*abs(l_ref.item())
    if use_automask: assert (ld_our['automask'].cpu() != ld_ref['automask']).float().mean().item() < 5e-3
    e = ((pc.grad.double().cpu() - p64.grad).norm()/p64.grad.norm()).item()
    assert e < (2e-2 if use_min or use_automask else 1e-4), e   # decisions flip at float32 near-ties; smooth case: 1e-4
