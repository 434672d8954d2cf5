"""GPU: libstv loss kernels (through the C ABI) vs the float64 oracle and the reference's golden fixtures.

Tolerances (float32 kernels vs float64 oracle, norm-wise relative error ||a-b||/||b||):
  * loss values              <= 1e-5
  * d loss / d disparity     <= 1e-4      (given identical per-pixel decisions, see below)
  * d loss / d(aa, t, K)     <= 1e-4      (idem; two-stage fixed-order reductions, double-precision final sum)
Discrete per-pixel decisions (which support frame wins the min-reprojection, auto-mask on/off) can legitimately flip
between two float32 evaluations when the competing errors differ by less than float32 resolution — the reference's own
float32 run differs from its float64 run by ~1e-2 on d/d disparity for this reason (see test_oracle_golden). The check is
therefore split: (1) decisions equal the oracle's except where the oracle's margin is < 2e-6 (and on < 0.5% of
pixels); (2) with the oracle forced to the kernel's decisions, values and gradients agree to the tolerances above.
"""
import numpy as np
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu

TOL_LOSS, TOL_GRAD = 1e-5, 1e-4


def _decision_margin(inp, cfg, o64):
    """Per-pixel margin of the oracle's decision: gap between the best and second-best candidate error."""
    from oracle import loss as OL
    d = U.cast(inp, torch.float64)
    n = d['supp_imgs'].shape[0]
    S, b = cfg['S'], cfg['b']
    # Candidate errors per (S*b, n [+1], H, W), recomputed from the oracle's pieces.
    H, W = d['imgs'].shape[-2:]
    Ts = OL.T_from_AAt(d['aa'], d['t'])
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    dep = torch.cat([OL.disp_to_depth(OL.resize_bilinear(x, (H, W)), mn, mx) for x in d['disps']], 0)
    tgt = d['imgs'].repeat(S, 1, 1, 1)
    fn = OL.photo_error if cfg.get('loss_name', 'ssim') == 'ssim' else (lambda p, t: (p - t).abs().mean(1, keepdim=True))
    errs = []
    for k in range(n):
        w, _, _ = OL.view_synth(d['supp_imgs'][k].repeat(S, 1, 1, 1), dep, Ts[k].repeat(S, 1, 1), d['K'].repeat(S, 1, 1))
        errs.append(fn(w, tgt))
    errs = torch.cat(errs, 1)
    cands = errs if cfg.get('use_min', True) else errs.mean(1, keepdim=True)
    if cfg.get('use_automask', True):
        st = OL.compute_photo(d['supp_imgs'].repeat(1, S, 1, 1, 1), tgt, cfg.get('use_min', True), cfg.get('loss_name', 'ssim'))
        st = st + torch.finfo(torch.float32).eps*d['noise']  # the kernels add float32 eps, like the reference in float32
        cands = torch.cat([cands, st], 1)
    if cands.shape[1] == 1: return torch.full_like(cands, float('inf'))
    top2 = cands.topk(2, dim=1, largest=False)[0]
    return (top2[:, 1:2] - top2[:, 0:1])


@pytest.mark.parametrize('name', U.LOSS_CASES)
def test_loss_stack_matches_oracle(name):
    inp, cfg, ref = U.load_golden(name)
    got = U.run_cuda(inp, cfg)
    torch.cuda.synchronize()
    sel = got['sel'].cpu()

    # (1) decisions
    free = U.run_oracle(inp, cfg, torch.float64)
    if cfg.get('use_min', True) or cfg.get('use_automask', True):
        osel = free['sel']
        if not cfg.get('use_min', True): osel = torch.where(osel == 255, osel, torch.full_like(osel, 254))
        diff = sel != osel
        margin = _decision_margin(inp, cfg, free)
        assert diff.float().mean().item() < 5e-3, f'{diff.float().mean().item():.4%} decisions differ'
        assert (margin[diff] < 2e-6).all(), f'decision flipped with margin {margin[diff].max().item():.3e}'

    # (2) values and gradients given the kernel's decisions
    fsel = sel if cfg.get('use_min', True) else None
    if fsel is None and cfg.get('use_automask', True):
        pytest.skip('mean-reduction with automask: forced decisions not defined in the oracle; covered by decisions + golden')
    want = U.run_oracle(inp, cfg, torch.float64, forced_sel=fsel)
    assert U.rel(got['loss_recon'], want['loss_recon']) < TOL_LOSS
    assert U.rel(got['loss_smooth'], want['loss_smooth']) < TOL_LOSS
    for s in range(cfg['S']):
        assert U.rel(got[f'g_disp{s}'], want[f'g_disp{s}']) < TOL_GRAD, f'g_disp{s}: {U.rel(got[f"g_disp{s}"], want[f"g_disp{s}"]):.3e}'
    for k in ('g_aa', 'g_t', 'g_K'):
        assert U.rel(got[k], want[k]) < TOL_GRAD, f'{k}: {U.rel(got[k], want[k]):.3e}'
    for k in ('warp0', 'depth_up0', 'disp_grad', 'image_grad'):
        assert U.rel(got[k], want[k]) < 1e-5, k


@pytest.mark.parametrize('name', U.LOSS_CASES)
def test_loss_values_match_reference_golden(name):
    """Directly against the reference's stored float64 answers (decisions included): losses to 1e-5, logging maps to 1e-5."""
    inp, cfg, ref = U.load_golden(name)
    got = U.run_cuda(inp, cfg)
    assert abs(got['loss_recon'].item() - ref['ref64_loss_recon'].item()) < 2e-5*abs(ref['ref64_loss_recon'].item())
    assert abs(got['loss_smooth'].item() - ref['ref64_loss_smooth'].item()) < 1e-5*abs(ref['ref64_loss_smooth'].item())
    for k in ('warp0', 'depth_up0', 'disp_grad', 'image_grad'):
        assert U.rel(got[k][..., ::4, ::4], torch.from_numpy(ref[f'ref64_{k}'])) < 1e-5, k
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
    feat = torch.rand(cfg['b'], 5, H, W, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
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
