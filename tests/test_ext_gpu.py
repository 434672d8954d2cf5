"""GPU: the 8f-4 row (other registered losses) through the C ABI against the REFERENCE's own float64 outputs (tests/golden/ext_cases.npz,
oracle/make_golden_ext.py) and the float64 oracle: ReconstructionLoss with weighting masks / 'l2' / C-channel inputs
(stv_recon_ex_fwd/bwd), SmoothReg with use_laplacian / use_blur (stv_smooth_ex_fwd/bwd), RegressionLoss (stv_regr_fwd/bwd), handlers.feat_recon / stereo_const / depth_regr, and
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


@pytest.mark.parametrize('S', [1, 2])
def test_image_recon_general_route_with_predicted_masks(S):
    """handlers.image_recon with `masks` (explainability): stand-alone warp + general loss vs the oracle's composition; S = 2 checks
    the (n, S, b) expansion order of handlers.py:43-66 (scale-major batches, masks stacked per scale)."""
    c = dict(b=2, n=2, C=3, H=16, W=24, seed=51)
    d = _cuda(G.feat_inputs(c, torch.float32))
    rs = np.random.RandomState(52)
    f = lambda *sh: torch.from_numpy(rs.random_sample(sh).astype(np.float32)).cuda()
    imgs = f(c['b'], 3, c['H'], c['W'])
    supp = (imgs[None] + 0.2*f(c['n'], c['b'], 3, c['H'], c['W']) - 0.1).clamp(0, 1)
    masks = {s: (0.2 + 0.6*f(c['b'], c['n'], c['H'], c['W'])).requires_grad_() for s in range(S)}
    depths = {s: (d['depth']*(1 + 0.25*s)).detach().requires_grad_() for s in range(S)}
    Ts = T_from_AAt(d['aa'], d['t'])
    crit = ReconstructionLoss('ssim', True, False, 'explainability')
    loss, ld = Hd.image_recon(crit, None, depths, masks, imgs, supp, Ts, d['K'])
    loss.backward()
    rdepth = [v.detach().double().cpu().requires_grad_() for v in depths.values()]
    rmask = [v.detach().double().cpu().requires_grad_() for v in masks.values()]
    rT, rK, rsupp, rimgs = Ts.double().cpu(), d['K'].double().cpu(), supp.double().cpu(), imgs.double().cpu()
    # the reference averages over the (S*b) virtual batch: equal-sized scales -> the mean of the per-scale losses
    want = 0
    for s in range(S):
        warp = torch.stack([O.view_synth(rsupp[k], rdepth[s], rT[k], rK)[0] for k in range(c['n'])])
        want = want + O.reconstruction_loss_ex(warp, rimgs, None, rmask[s], 'ssim', True, False, 'explainability')[0]/S
    want.backward()
    assert abs(loss.item() - want.item()) <= 2e-5*abs(want.item())
    for s in range(S):
        assert U.rel(masks[s].grad.cpu().double(), rmask[s].grad) < TOL_GRAD
        assert U.rel(depths[s].grad.cpu().double(), rdepth[s].grad) < 2e-3   # min-reprojection flips at float32 near-ties stay local
    assert ld['supp_imgs_warp'].shape == supp.shape


@pytest.mark.parametrize('name', list(G.REGR))
def test_regression_loss_matches_the_reference_class(name):
    from slowtv_monodepth_b200.losses import RegressionLoss
    c = G.REGR[name]
    d = _cuda(G.regr_inputs(c, torch.float32))
    pred, tgt = d['pred'].requires_grad_(), d['tgt'].requires_grad_()
    loss, ld = RegressionLoss(c['loss_name'], invert=c['invert'])(pred, tgt, d['mask'] if c['masked'] else None)
    loss.backward()
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= TOL_LOSS*abs(want)
    assert U.rel(pred.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_pred'])) < TOL_GRAD
    assert U.rel(tgt.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_tgt'])) < TOL_GRAD
    assert U.rel(ld['err_regr'].cpu().double(), torch.from_numpy(EXT[f'{name}/err'])) < 1e-5


def test_regression_loss_is_deterministic_and_splits_ties():
    """Two runs give bit-identical results (no atomics); berHu's threshold gradient is shared evenly between tied maxima."""
    from slowtv_monodepth_b200 import functional as F_
    rs = np.random.RandomState(7)
    pred = torch.from_numpy(rs.random_sample((3, 1, 33, 47)).astype(np.float32)).cuda()
    tgt = torch.zeros_like(pred)
    pred.view(-1)[5] = pred.view(-1)[900] = 4.0   # two tied maxima of |pred - target|
    outs = []
    for _ in range(2):
        p = pred.clone().requires_grad_()
        loss, _ = F_.regr_loss(p, tgt, None, loss_name='berhu')
        loss.backward()
        outs.append((loss.clone(), p.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    pr = pred.double().cpu().requires_grad_()
    want = O.regression_loss(pr, tgt.double().cpu(), None, 'berhu')[0]
    want.backward()
    assert abs(outs[0][0].item() - want.item()) <= TOL_LOSS*abs(want.item())
    assert U.rel(outs[0][1].cpu().double(), pr.grad) < TOL_GRAD


@pytest.mark.parametrize('name', list(G.STEREO))
def test_stereo_const_matches_the_reference_handler(name):
    from slowtv_monodepth_b200.geometry import to_scaled
    from slowtv_monodepth_b200.losses import RegressionLoss
    c = G.STEREO[name]
    d = G.stereo_inputs(c, torch.float32)
    disps = {s: x.cuda().requires_grad_() for s, x in enumerate(d['disps'])}
    disps_st = {s: x.cuda().requires_grad_() for s, x in enumerate(d['disps_st'])}
    depths = {s: to_scaled(x, 0.1, 100.)[1] for s, x in disps.items()}
    depths_st = {s: to_scaled(x, 0.1, 100.)[1] for s, x in disps_st.items()}
    loss, ld = Hd.stereo_const(RegressionLoss(c['loss_name']), None, disps, depths, disps_st, depths_st, d['T'].cuda(), d['K'].cuda())
    loss.backward()
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= 2e-5*abs(want)
    for s in disps:
        assert U.rel(disps[s].grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_disp{s}'])) < 5e-4      # through the float32 warp + sampler
        assert U.rel(disps_st[s].grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_disp_st{s}'])) < 5e-4
    assert U.rel(ld['disps_warp'].cpu().double(), torch.from_numpy(EXT[f'{name}/disps_warp']).double()) < 1e-4


@pytest.mark.parametrize('name', list(G.HINTS))
def test_depth_regr_matches_the_reference_handler(name):
    from slowtv_monodepth_b200.losses import RegressionLoss
    c = G.HINTS[name]
    d = G.hints_inputs(c, torch.float32)
    depths = {s: x.cuda().requires_grad_() for s, x in enumerate(d['depths'])}
    photo = ReconstructionLoss('ssim', use_min=True).compute_photo
    crit = RegressionLoss(c['loss_name'], invert=c['invert'], use_automask=c['use_automask'])
    loss, ld = Hd.depth_regr(crit, None, photo, depths, d['hints'].cuda(), d['imgs'].cuda(), d['supp'].cuda(),
                             T_from_AAt(d['aa'].cuda(), d['t'].cuda()), d['K'].cuda())
    loss.backward()
    got_mask = ld['mask_regr'].cpu().numpy().astype(np.uint8)
    flips = (got_mask != EXT[f'{name}/mask']).mean()
    assert flips <= 0.01                                              # automask comparisons at float32 near-ties
    if flips: pytest.skip('an automask decision flipped at a float32 near-tie: loss / gradients not comparable')
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= 2e-5*abs(want)
    for s in depths: assert U.rel(depths[s].grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_depth{s}'])) < TOL_GRAD


@pytest.mark.parametrize('shape', [(1, 1, 3, 3, 3), (2, 1, 4, 3, 5), (1, 3, 2, 5, 4)])   # (b, n, C, H, W): minimum sizes, single frame, odd shapes
@pytest.mark.parametrize('loss_name,mask_name', [('ssim', 'uncertainty'), ('l2', None), ('l1', 'explainability')])
def test_recon_ex_edge_shapes_match_oracle(shape, loss_name, mask_name):
    """Smallest legal maps (3 x 3: every pixel is a reflected border), one frame handed as a 4-D tensor, ragged sizes."""
    b, n, C, H, W = shape
    d = _cuda(G.recon_inputs(dict(b=b, n=n, C=C, H=H, W=W, seed=90 + H*W), torch.float32))
    pred = (d['pred'][0] if n == 1 else d['pred']).contiguous().requires_grad_()
    src = (d['src'][0] if n == 1 else d['src']).contiguous()
    mask = d['mask'].requires_grad_()
    crit = ReconstructionLoss(loss_name, True, True, mask_name)
    loss, ld = crit(pred, d['tgt'], source=src, mask=mask if mask_name else None, noise=d['noise'])
    loss.backward()
    rp, rm = d['pred'].detach().double().cpu().requires_grad_(), d['mask'].detach().double().cpu().requires_grad_()
    want, automask, _, _ = O.reconstruction_loss_ex(rp, d['tgt'].double().cpu(), d['src'].double().cpu(), rm if mask_name else None, loss_name, True, True,
                                                    mask_name, d['noise'].double().cpu())
    want.backward()
    if not torch.equal(ld['automask'].cpu(), automask): pytest.skip('a decision flipped at a float32 near-tie')
    assert abs(loss.item() - want.item()) <= TOL_LOSS*abs(want.item())
    assert U.rel(pred.grad.cpu().double().reshape(rp.shape), rp.grad) < TOL_GRAD
    if mask_name: assert U.rel(mask.grad.cpu().double(), rm.grad) < TOL_GRAD


@pytest.mark.parametrize('flags', [(True, True, True), (False, False, True), (True, True, False)])
def test_smooth_ex_minimum_size(flags):
    d = _cuda(G.smooth_inputs(dict(b=2, H=3, W=4, seed=95), torch.float32))
    disp = d['disp'].requires_grad_()
    loss, _ = SmoothReg(*flags)(disp, d['img'])
    loss.backward()
    rd = d['disp'].detach().double().cpu().requires_grad_()
    want = O.smooth_reg_ex(rd, d['img'].double().cpu(), *flags)[0]
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_LOSS*abs(want.item())
    assert U.rel(disp.grad.cpu().double(), rd.grad) < TOL_GRAD


def test_regression_loss_edge_cases():
    """All-zero mask rows, zero targets under `invert` (to_inv's guard), identical inputs (sign(0) = 0, berHu threshold 0)."""
    from slowtv_monodepth_b200 import functional as F_
    rs = np.random.RandomState(3)
    pred = torch.from_numpy((0.1 + rs.random_sample((2, 1, 5, 7))).astype(np.float32)).cuda()
    tgt = pred.clone(); tgt[0, 0, :2] = 0.0
    mask = (tgt > 0).float()
    for name in ('l1', 'log_l1', 'berhu'):
        p = pred.clone().requires_grad_()
        loss, err = F_.regr_loss(p, tgt, mask, loss_name=name, invert=True)
        loss.backward()
        pr = pred.double().cpu().requires_grad_()
        want = O.regression_loss(pr, tgt.double().cpu(), mask.double().cpu(), name, True)[0]
        want.backward()
        assert abs(loss.item() - want.item()) <= 1e-6 + TOL_LOSS*abs(want.item())
        assert (p.grad.cpu().double() - pr.grad).abs().max().item() <= 1e-6 + TOL_GRAD*pr.grad.abs().max().item()


@pytest.mark.parametrize('name', list(G.FREG))
def test_feature_regularisers_match_the_reference_classes(name):
    from slowtv_monodepth_b200 import regularizers as R
    c = G.FREG[name]
    d = _cuda(G.freg_inputs(c, torch.float32))
    feat = d['feat'].requires_grad_()
    loss, ld = getattr(R, c['cls'])(use_edges=c['use_edges'])(feat, d['img'])
    loss.backward()
    want = EXT[f'{name}/loss'].item()
    assert abs(loss.item() - want) <= TOL_LOSS*abs(want)
    assert U.rel(feat.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_feat'])) < TOL_GRAD
    assert U.rel(ld['feat_grad'].cpu().double(), torch.from_numpy(EXT[f'{name}/feat_grad']).clamp(min=O.EPS32**0.5)) < 1e-5


def test_pointwise_regularisers_and_their_handlers():
    from slowtv_monodepth_b200 import regularizers as R
    d = _cuda(G.freg_inputs(G.FREG['freg_peaky_edges'], torch.float32))
    for name, crit in (('pw_mask', R.MaskReg()), ('pw_occ', R.OccReg()), ('pw_occ_inv', R.OccReg(invert=True))):
        x = d['x'].clone().requires_grad_()
        loss, ld = crit(x)
        loss.backward()
        want = EXT[f'{name}/loss'].item()
        assert ld == {} and abs(loss.item() - want) <= TOL_LOSS*abs(want)
        assert U.rel(x.grad.cpu().double(), torch.from_numpy(EXT[f'{name}/g_x'])) < TOL_GRAD
    maps = {0: d['x'].clone().requires_grad_(), 1: d['x'][..., ::2, ::2].clone().requires_grad_()}
    loss, _ = Hd.disp_occ(R.OccReg(), maps)
    loss.backward()
    assert abs(loss.item() - 0.5*(maps[0].mean().item() + maps[1].mean().item())) < 1e-6
    loss, _ = Hd.disp_mask(R.MaskReg(), {0: d['x'].clone()})
    assert abs(loss.item() - EXT['pw_mask/loss'].item()) <= TOL_LOSS*abs(EXT['pw_mask/loss'].item())


def test_feat_smooth_handler_follows_the_reference_formulation():
    from slowtv_monodepth_b200 import regularizers as R
    rs = np.random.RandomState(5)
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).cuda()
    imgs, supp = f(2, 3, 16, 24), f(2, 2, 3, 16, 24)
    feats = [f(2, 4, 16, 24).requires_grad_(), f(2, 6, 8, 12).requires_grad_()]
    sfeats = [f(2, 2, 4, 16, 24).requires_grad_(), f(2, 2, 6, 8, 12).requires_grad_()]
    loss, _ = Hd.feat_smooth(R.FeatSmoothReg(use_edges=True), feats, imgs, sfeats, supp)
    loss.backward()
    rf = [x.detach().double().cpu().requires_grad_() for x in feats]
    rsf = [x.detach().double().cpu().requires_grad_() for x in sfeats]
    im, sim = imgs.double().cpu(), supp.double().cpu().flatten(0, 1)
    one = lambda fs, img: torch.stack([O.feat_smooth_reg(x, O.resize_bilinear(img, x.shape[-2:]), True)[0]/2**s for s, x in enumerate(fs)]).mean()
    want = one(rf, im) + one([x.flatten(0, 1) for x in rsf], sim)
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_LOSS*abs(want.item())
    for a, b in zip(feats + sfeats, rf + rsf): assert U.rel(a.grad.cpu().double(), b.grad) < TOL_GRAD
