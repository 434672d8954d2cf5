"""GPU: libstv loss kernels (through the C ABI) vs the float64 oracle and the reference's golden fixtures.

Tolerances (float32 kernels vs the float64 oracle, norm-wise relative error ||a-b||/||b||):
  * loss values                         <= 1e-5
  * d loss / d disparity, d/d depth_up  <= 1e-4   on stable pixels, given identical per-pixel decisions (see below)
  * d loss / d(aa, t, K)                <= 1e-4   (two-stage fixed-order reductions, double-precision final sum)

The loss is only piecewise smooth. Two float32 evaluations (ours, or the reference's own) can legitimately take a
different branch at a pixel when a discrete event sits within float32 rounding of flipping:
  (a) decisions — which support frame wins the min-reprojection, auto-mask on/off (torch.min over candidates);
  (b) sub-gradient events — sign(warp - target) of the L1 term, the texel cell floor(ix) of the bilinear sampler, the
      border / depth clamps.
The reference's float32 run differs from its float64 run by ~1e-2 on d/d disparity for exactly this reason
(tests/test_oracle_golden.py::test_fp32_reference_noise_floor), so the check is split:
  (1) decisions equal the oracle's except where the oracle's margin is < 2e-5, and on < 0.5 % of the pixels;
  (2) the oracle is re-run with the kernel's decisions forced (oracle/loss.py `forced_sel`); values and pose/intrinsics
      gradients must then agree to the tolerances above; the full-resolution per-pixel gradient maps (d/d depth_up) are compared
      on the pixels whose 3x3 neighbourhood holds no type-(b) event (tests/util.py::unstable_pixels) — at least 85 % of every
      scale's pixels must take part (asserted) — and the low-resolution d/d disparity on EVERY pixel, against the oracle's own
      vector-Jacobian product of the bilinear upsample + to_scaled applied to the kernel's full-resolution maps plus the
      oracle's smoothness gradient (both event-free), tests/util.py::check_pixel_gradients.
The oracle runs with float32's eps (`util.eps32`), i.e. it is the exact-arithmetic version of the float32 reference.
"""
import numpy as np
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu

TOL_LOSS, TOL_GRAD = 1e-5, 1e-4


@pytest.mark.parametrize('mode', ['fused-disp', 'fused-depth', 'two-pass'])
@pytest.mark.parametrize('name', U.LOSS_CASES)
def test_loss_stack_matches_oracle(name, mode):
    """All three routes: the single-pass kernel fed with low-resolution disparities (the training step) or with up-sampled depth
    maps (the reference's handler signature), and the two-pass kernels (always used for the mean reduction)."""
    inp, cfg, ref = U.load_golden(name)
    if mode != 'two-pass' and not cfg.get('use_min', True): pytest.skip('the single-pass kernel needs min-reprojection; covered by two-pass')
    got = U.run_cuda(inp, cfg, mode=mode)
    torch.cuda.synchronize()
    U.check_loss_stack(inp, cfg, got, TOL_LOSS, TOL_GRAD)


@pytest.mark.parametrize('name', U.LOSS_CASES)
def test_loss_values_match_reference_golden(name):
    """Directly against the reference's stored float64 answers (decisions included): losses to 1e-5, logging maps to 1e-5."""
    inp, cfg, ref = U.load_golden(name)
    got = U.run_cuda(inp, cfg)
    assert abs(got['loss_recon'].item() - ref['ref64_loss_recon'].item()) < 2e-5*abs(ref['ref64_loss_recon'].item())
    assert abs(got['loss_smooth'].item() - ref['ref64_loss_smooth'].item()) < 1e-5*abs(ref['ref64_loss_smooth'].item())
    for k in ('warp0', 'depth_up0', 'disp_grad', 'image_grad'):
        r = torch.from_numpy(ref[f'ref64_{k}']).double()
        live = (r > 1e-3).double()  # the stored float64 maps bottom out at sqrt(eps64); the float32 path at sqrt(eps32)
        assert U.rel(got[k][..., ::4, ::4].cpu().double()*live, r*live) < 1e-5, k
    if 'ref64_automask0' in ref:
        mism = (got['automask0'].cpu().numpy().astype(np.uint8) != ref['ref64_automask0']).mean()
        assert mism < 5e-3, f'automask differs on {mism:.3%} of pixels'


def test_mean_automask_against_golden_gradients():
    """use_min=False + automask (no forced-decision mode): compare gradients with the reference directly, loosely."""
    inp, cfg, ref = U.load_golden('noscale_mean_auto')
    got = U.run_cuda(inp, cfg)
    for k in ('g_aa', 'g_t', 'g_K'):
        assert U.rel(got[k], torch.from_numpy(ref[f'ref64_{k}'])) < 2e-2, k


def test_deterministic():
    inp, cfg, _ = U.load_golden('ragged_n4')
    a, b = U.run_cuda(inp, cfg), U.run_cuda(inp, cfg)
    for k in ('loss_recon', 'loss_smooth', 'g_aa', 'g_t', 'g_K', 'g_disp0', 'g_disp1'):
        assert torch.equal(a[k], b[k]), k


def test_view_synth_module_matches_oracle():
    from oracle import loss as OL
    from slowtv_monodepth_b200.geometry import ViewSynth
    inp, cfg, _ = U.load_golden('ragged_n4')
    d64 = U.cast(inp, torch.float64)
    H, W = cfg['shape']
    # Smooth features: with white-noise features d/dT is a sum of ~H*W terms of random sign (condition number ~1e3), which
    # no float32 evaluation — the reference's included — resolves to 1e-4.
    g = torch.Generator().manual_seed(0)
    feat = torch.nn.functional.interpolate(torch.rand(cfg['b'], 5, 4, 6, dtype=torch.float64, generator=g), size=(H, W),
                                           mode='bicubic', align_corners=True)
    depth = OL.disp_to_depth(d64['disps'][0], 0.1, 100.)
    T = OL.T_from_AAt(d64['aa'][0], d64['t'][0])

    def run(fn, dt, dev):
        x, dp, Tm, K = (v.to(dev, dt).clone().requires_grad_() for v in (feat, depth, T, d64['K']))
        w, dw, valid = fn(x, dp, Tm, K)
        ((w*w).sum() + dw.sum()).backward()
        return w.detach(), dw.detach(), valid, x.grad, dp.grad, Tm.grad, K.grad

    want = run(OL.view_synth, torch.float64, 'cpu')
    got = run(ViewSynth((H, W)), torch.float32, 'cuda')
    assert U.rel(got[0], want[0]) < 1e-5 and U.rel(got[1], want[1]) < 1e-6
    assert (got[2].cpu() != want[2]).float().mean() < 1e-3
    for j, name in ((3, 'g_input'), (4, 'g_depth'), (5, 'g_T'), (6, 'g_K')):
        assert U.rel(got[j], want[j]) < 1e-4, f'{name}: {U.rel(got[j], want[j]):.3e}'


def test_argument_errors_are_value_errors():
    from slowtv_monodepth_b200 import functional as F_
    t = torch.zeros(1, 3, 8, 8, device='cuda')
    with pytest.raises(ValueError): F_.photo_loss([torch.zeros(1, 1, 8, 9, device='cuda')], t, t[None], torch.eye(4, device='cuda')[None, None], torch.eye(4, device='cuda')[None])
    with pytest.raises(ValueError): F_.disp_to_depth(torch.zeros(1, 1, 4, 4, device='cuda'), (8, 8), -1.0, 100.)
    with pytest.raises(Exception): F_.photo_loss([torch.zeros(1, 1, 8, 8)], t.cpu(), t[None].cpu(), torch.eye(4)[None, None], torch.eye(4)[None])
