"""GPU: the 8f-4 row (other registered losses) through the C ABI against the REFERENCE's own float64 outputs (tests/golden/ext_cases.npz,
oracle/make_golden_ext.py) and the float64 oracle: ReconstructionLoss with weighting masks / 'l2' / C-channel inputs
(stv_recon_ex_fwd/bwd), SmoothReg with use_laplacian / use_blur (stv_smooth_ex_fwd/bwd), handlers.feat_recon, and
handlers.image_recon's general route with predicted masks. Tolerances: loss 1e-5 rel, gradients 1e-4 rel (float32 kernels)."""
import numpy as np
import pytest
import torch

from oracle import loss as O
from oracle import make_golden_ext as G
from slowtv_monodepth_b200 import handlers as Hd
from slowtv_monodepth_b200.geometry import T_from_AAt
from slowtv_monodepth_b200.losses import ReconstructionLoss
from slowtv_monodepth_b200.regularizers import SmoothReg
from tests import util as U

pytestmark = pytest.mark.gpu
EXT = np.load(U.GOLDEN/'ext_cases.npz')
TOL_LOSS, TOL_GRAD = 1e-5, 1e-4


def _cuda(d):
    return {k: v.to(torch.float32).cuda() for k, v in d.items()}


@pytest.mark.parametrize('name', list(G.RECON))
def test_recon_ex_matches_the_reference_class(name):
    c = G.RECON[name]
    d = _cuda(G.recon_inputs(c, torch.float32))
    pred, mask = d['pred'].requires_grad_(), d['mask'].requires_grad_()
    crit = ReconstructionLoss(c['loss_name'], c['use_min'], c['use_automask'], c['mask_name'])
    loss, ld = crit(pred, d['tgt'], source=d['src'], mask=mask if c['mask_name'] else None, noise=d['noise'] if c['use_automask'] else None)
    loss.backward()
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= TOL_LOSS*abs(want)
    if c['use_automask']:
        got = ld['automask'].cpu().numpy().astype(np.uint8)
        assert (got != EXT[f'{name}/automask']).mean() <= 0.005   # float32 vs float64 near-ties
        if (got != EXT[f'{name}/automask']).any(): pytest.skip('a decision flipped at a float32 near-tie: gradients not comparable')
    assert U.rel(pred.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_pred'])) < TOL_GRAD
    if c['mask_name']: assert U.rel(mask.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_mask'])) < TOL_GRAD


def test_recon_ex_draws_its_own_noise_and_counts_calls():
    c = G.RECON['recon_ssim_min_auto_expl']
    d = _cuda(G.recon_inputs(c, torch.float32))
    crit = ReconstructionLoss(c['loss_name'], c['use_min'], c['use_automask'], c['mask_name'])
    l1, _ = crit(d['pred'], d['tgt'], source=d['src'], mask=d['mask'])
    l2, _ = crit(d['pred'], d['tgt'], source=d['src'], mask=d['mask'])
    assert crit.noise_step.item() == 2 and abs(l1.item() - l2.item()) < 1e-4


def test_broadcast_mask_and_errors():
    c = G.RECON['recon_l1_min_unc']
    d = _cuda(G.recon_inputs(c, torch.float32))
    crit = ReconstructionLoss('l1', True, False, 'uncertainty')
    m1 = d['mask'][:, :1].contiguous().requires_grad_()
    loss, _ = crit(d['pred'], d['tgt'], mask=m1)
    loss.backward()
    ref_m = d['mask'][:, :1].double().cpu().expand(-1, c['n'], -1, -1).clone().requires_grad_()
    want = O.reconstruction_loss_ex(d['pred'].double().cpu(), d['tgt'].double().cpu(), None, ref_m, 'l1', True, False, 'uncertainty')[0]
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_LOSS*abs(want.item())
    assert U.rel(m1.grad.cpu().double(), ref_m.grad.sum(1, keepdim=True)) < TOL_GRAD
    with pytest.raises(ValueError): crit(d['pred'], d['tgt'])                       # mask_name without a mask
    with pytest.raises(ValueError): ReconstructionLoss('ssim', mask_name='nope')
    with pytest.raises(KeyError): ReconstructionLoss('huber')


@pytest.mark.parametrize('name', list(G.SMOOTH))
def test_smooth_ex_matches_the_reference_class(name):
    c = G.SMOOTH[name]
    d = _cuda(G.smooth_inputs(c, torch.float32))
    disp = d['disp'].requires_grad_()
    loss, ld = SmoothReg(use_edges=c['use_edges'], use_laplacian=c['use_laplacian'])(disp, d['img'])   # first order: the hot-path kernels
    loss.backward()
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= TOL_LOSS*abs(want)
    assert U.rel(disp.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_disp'])) < TOL_GRAD
    # the logging maps clamp at eps(dtype) before the square root (smooth.py:89,92): float32's eps on this side, float64's in the fixture
    for key in ('disp_grad', 'image_grad'):
        want_map = torch.from_numpy(EXT[f'{name}/{key}']).clamp(min=O.EPS32**0.5)
        assert U.rel(ld[key].cpu().double(), want_map) < 1e-5, key


@pytest.mark.parametrize('flags', [(True, False, True), (True, True, True), (False, True, True)])
def test_smooth_ex_blur_matches_oracle(flags):
    """kornia is absent: the blur is checked against the oracle's restatement (parity unpinned)."""
    use_edges, use_lap, use_blur = flags
    d = _cuda(G.smooth_inputs(dict(b=2, H=13, W=17, seed=41), torch.float32))
    disp = d['disp'].requires_grad_()
    loss, ld = SmoothReg(use_edges, use_lap, use_blur)(disp, d['img'])
    loss.backward()
    rd = d['disp'].detach().double().cpu().requires_grad_()
    want, dg, ig = O.smooth_reg_ex(rd, d['img'].double().cpu(), use_edges, use_lap, use_blur)
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_LOSS*abs(want.item())
    assert U.rel(disp.grad.cpu().double(), rd.grad) < TOL_GRAD
    lo = O.EPS32**0.5
    assert U.rel(ld['disp_grad'].cpu().double(), dg.detach().clamp(min=lo)) < 1e-5 and U.rel(ld['image_grad'].cpu().double(), ig.clamp(min=lo)) < 1e-5


def test_general_multi_scale_smoothness_follows_the_handler():
    d = _cuda(G.smooth_inputs(dict(b=2, H=16, W=24, seed=42), torch.float32))
    disps = {0: d['disp'].clone().requires_grad_(), 1: d['disp'][..., ::2, ::2].clone().requires_grad_()}
    loss, ld = Hd.disp_smooth(SmoothReg(True, True, False), disps, d['img'])
    loss.backward()
    rd = {k: v.detach().double().cpu().requires_grad_() for k, v in disps.items()}
    img = d['img'].double().cpu()
    want = torch.stack([O.smooth_reg_ex(v, O.resize_bilinear(img, v.shape[-2:]), True, True, False)[0]/2**k for k, v in rd.items()]).mean()
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_LOSS*abs(want.item())
    for k in disps: assert U.rel(disps[k].grad.cpu().double(), rd[k].grad) < TOL_GRAD
    assert ld['disp_grad'].shape == d['disp'].shape


@pytest.mark.parametrize('name', list(G.FEAT))
def test_feat_recon_matches_the_reference_handler(name):
    c = G.FEAT[name]
    d = _cuda(G.feat_inputs(c, torch.float32))
    depth = d['depth'].requires_grad_()
    crit = ReconstructionLoss(c['loss_name'], c['use_min'], c['use_automask'])
    if c['use_automask']:   # explicit noise has no door in the handler signature: reproduce the reference's draw through the seam below
        orig = crit.forward
        crit.forward = lambda *a, **k: orig(*a, **{**k, 'noise': d['noise']})
    loss, ld = Hd.feat_recon(crit, None, {0: depth}, None, d['feats'], d['supp'], T_from_AAt(d['aa'], d['t']), d['K'])
    loss.backward()
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= 2e-5*abs(want)
    assert U.rel(ld['supp_feats_warp'].cpu().double(), torch.from_numpy(EXT[f'{name}/warp']).double()) < 1e-4
    assert U.rel(depth.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_depth'])) < 5e-4   # through the float32 warp + sampler


def test_image_recon_general_route_with_predicted_masks():
    """handlers.image_recon with `masks` (explainability): stand-alone warp + general loss vs the oracle's composition."""
    c = dict(b=2, n=2, C=3, H=16, W=24, seed=51)
    d = _cuda(G.feat_inputs(c, torch.float32))
    rs = np.random.RandomState(52)
    imgs = torch.from_numpy(rs.random_sample((c['b'], 3, c['H'], c['W'])).astype(np.float32)).cuda()
    supp = (imgs[None] + 0.2*torch.from_numpy(rs.random_sample((c['n'], c['b'], 3, c['H'], c['W'])).astype(np.float32)).cuda() - 0.1).clamp(0, 1)
    mask = (0.2 + 0.6*torch.from_numpy(rs.random_sample((c['b'], c['n'], c['H'], c['W'])).astype(np.float32))).cuda().requires_grad_()
    depth = d['depth'].requires_grad_()
    Ts = T_from_AAt(d['aa'], d['t'])
    crit = ReconstructionLoss('ssim', True, False, 'explainability')
    loss, ld = Hd.image_recon(crit, None, {0: depth}, {0: mask}, imgs, supp, Ts, d['K'])
    loss.backward()
    rdepth, rmask = depth.detach().double().cpu().requires_grad_(), mask.detach().double().cpu().requires_grad_()
    rT, rK = Ts.double().cpu(), d['K'].double().cpu()
    warp = torch.stack([O.view_synth(supp[k].double().cpu(), rdepth, rT[k], rK)[0] for k in range(c['n'])])
    want = O.reconstruction_loss_ex(warp, imgs.double().cpu(), None, rmask, 'ssim', True, False, 'explainability')[0]
    want.backward()
    assert abs(loss.item() - want.item()) <= 2e-5*abs(want.item())
    assert U.rel(mask.grad.cpu().double(), rmask.grad) < TOL_GRAD
    assert U.rel(depth.grad.cpu().double(), rdepth.grad) < 2e-3   # min-reprojection flips at float32 near-ties stay local
    assert ld['supp_imgs_warp'].shape == supp.shape
