"""CPU: the oracle restatement of the 8f-4 row (oracle/loss.py: reconstruction_loss_ex, smooth_reg_ex, feat_recon) reproduces the
reference's own classes — fixtures in tests/golden/ext_cases.npz made by oracle/make_golden_ext.py from /root/reference, float64."""
import numpy as np
import pytest
import torch

from oracle import loss as O
from oracle import make_golden_ext as G
from tests import util as U

EXT = np.load(U.GOLDEN/'ext_cases.npz')


@pytest.mark.parametrize('name', list(G.RECON))
def test_recon_ex_matches_reference(name):
    c, d = G.RECON[name], G.recon_inputs(G.RECON[name])
    pred, mask = d['pred'].clone().requires_grad_(), d['mask'].clone().requires_grad_()
    loss, automask, _, sel = O.reconstruction_loss_ex(pred, d['tgt'], d['src'], mask if c['mask_name'] else None, c['loss_name'],
                                                      c['use_min'], c['use_automask'], c['mask_name'], d['noise'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-12*abs(EXT[f'{name}/loss'].item())
    assert U.rel(pred.grad, torch.from_numpy(EXT[f'{name}/g_pred'])) < 1e-10
    if c['mask_name']: assert U.rel(mask.grad, torch.from_numpy(EXT[f'{name}/g_mask'])) < 1e-10
    if c['use_automask']:
        assert np.array_equal(automask.numpy().astype(np.uint8), EXT[f'{name}/automask'])
        assert np.array_equal((sel.numpy() < 128), automask.numpy())


@pytest.mark.parametrize('name', list(G.SMOOTH))
def test_smooth_ex_matches_reference(name):
    c, d = G.SMOOTH[name], G.smooth_inputs(G.SMOOTH[name])
    disp = d['disp'].clone().requires_grad_()
    loss, dg, ig = O.smooth_reg_ex(disp, d['img'], c['use_edges'], c['use_laplacian'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-12*abs(EXT[f'{name}/loss'].item())
    assert U.rel(disp.grad, torch.from_numpy(EXT[f'{name}/g_disp'])) < 1e-10
    assert U.rel(dg, torch.from_numpy(EXT[f'{name}/disp_grad'])) < 1e-12 and U.rel(ig, torch.from_numpy(EXT[f'{name}/image_grad'])) < 1e-12


@pytest.mark.parametrize('name', list(G.FEAT))
def test_feat_recon_matches_reference(name):
    c, d = G.FEAT[name], G.feat_inputs(G.FEAT[name])
    depth = d['depth'].clone().requires_grad_()
    loss, warp = O.feat_recon(depth, d['feats'], d['supp'], O.T_from_AAt(d['aa'], d['t']), d['K'], c['loss_name'], c['use_min'],
                              c['use_automask'], noise=d['noise'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-10*abs(EXT[f'{name}/loss'].item())
    assert U.rel(depth.grad, torch.from_numpy(EXT[f'{name}/g_depth'])) < 1e-8
    assert U.rel(warp, torch.from_numpy(EXT[f'{name}/warp']).double()) < 1e-6   # stored as float32


def test_blur_is_a_normalised_reflecting_gaussian():
    """kornia is absent (parity unpinned): the restated 3x3 sigma-1 blur must at least keep constants and be symmetric."""
    x = torch.rand(1, 2, 7, 9, dtype=torch.float64)
    assert torch.allclose(O.gaussian_blur3(torch.ones_like(x)), torch.ones_like(x), atol=1e-15)
    assert torch.allclose(O.gaussian_blur3(x.flip(-1)).flip(-1), O.gaussian_blur3(x), atol=1e-15)
    assert O.gaussian_blur3(x).shape == x.shape


@pytest.mark.parametrize('name', list(G.REGR))
def test_regression_loss_matches_reference(name):
    c, d = G.REGR[name], G.regr_inputs(G.REGR[name])
    pred, tgt = d['pred'].clone().requires_grad_(), d['tgt'].clone().requires_grad_()
    loss, err = O.regression_loss(pred, tgt, d['mask'] if c['masked'] else None, c['loss_name'], c['invert'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-12*abs(EXT[f'{name}/loss'].item())
    assert U.rel(pred.grad, torch.from_numpy(EXT[f'{name}/g_pred'])) < 1e-10 and U.rel(tgt.grad, torch.from_numpy(EXT[f'{name}/g_tgt'])) < 1e-10
    assert U.rel(err, torch.from_numpy(EXT[f'{name}/err'])) < 1e-12


@pytest.mark.parametrize('name', list(G.STEREO))
def test_stereo_const_matches_reference(name):
    c, d = G.STEREO[name], G.stereo_inputs(G.STEREO[name])
    disps = [x.clone().requires_grad_() for x in d['disps']]
    disps_st = [x.clone().requires_grad_() for x in d['disps_st']]
    loss, warp = O.stereo_const(disps, [O.disp_to_depth(x, 0.1, 100.) for x in disps], disps_st, [O.disp_to_depth(x, 0.1, 100.) for x in disps_st],
                                d['T'], d['K'], c['loss_name'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-10*abs(EXT[f'{name}/loss'].item())
    for s in range(c['S']):
        assert U.rel(disps[s].grad, torch.from_numpy(EXT[f'{name}/g_disp{s}'])) < 1e-8
        assert U.rel(disps_st[s].grad, torch.from_numpy(EXT[f'{name}/g_disp_st{s}'])) < 1e-8
    first = warp.chunk(2)[1][:c['b']]   # `disps_warp` of the first scale
    assert U.rel(first, torch.from_numpy(EXT[f'{name}/disps_warp']).double()) < 1e-6


@pytest.mark.parametrize('name', list(G.HINTS))
def test_depth_regr_matches_reference(name):
    c, d = G.HINTS[name], G.hints_inputs(G.HINTS[name])
    depths = [x.clone().requires_grad_() for x in d['depths']]
    loss, masks = O.depth_regr(depths, d['hints'], d['imgs'], d['supp'], O.T_from_AAt(d['aa'], d['t']), d['K'], c['loss_name'], c['invert'],
                               c['use_automask'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-10*abs(EXT[f'{name}/loss'].item())
    for s in range(c['S']): assert U.rel(depths[s].grad, torch.from_numpy(EXT[f'{name}/g_depth{s}'])) < 1e-8
    assert np.array_equal(masks[:c['b']].numpy().astype(np.uint8), EXT[f'{name}/mask'])


@pytest.mark.parametrize('name', list(G.FREG))
def test_feature_regularisers_match_reference(name):
    c, d = G.FREG[name], G.freg_inputs(G.FREG[name])
    feat = d['feat'].clone().requires_grad_()
    fn = O.feat_peak_reg if c['cls'] == 'FeatPeakReg' else O.feat_smooth_reg
    loss, fg = fn(feat, d['img'], c['use_edges'])
    loss.backward()
    assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-12*abs(EXT[f'{name}/loss'].item())
    assert U.rel(feat.grad, torch.from_numpy(EXT[f'{name}/g_feat'])) < 1e-10 and U.rel(fg, torch.from_numpy(EXT[f'{name}/feat_grad'])) < 1e-12


def test_pointwise_regularisers_match_reference():
    d = G.freg_inputs(G.FREG['freg_peaky_edges'])
    for name, fn in (('pw_mask', O.mask_reg), ('pw_occ', O.occ_reg), ('pw_occ_inv', lambda x: O.occ_reg(x, True))):
        x = d['x'].clone().requires_grad_()
        loss = fn(x)
        loss.backward()
        assert abs(loss.item() - EXT[f'{name}/loss'].item()) <= 1e-12*abs(EXT[f'{name}/loss'].item())
        assert U.rel(x.grad, torch.from_numpy(EXT[f'{name}/g_x'])) < 1e-12
