"""Golden fixtures for the 8f-4 row (other registered losses): the REAL reference classes executed on CPU in float64.

TEST INFRASTRUCTURE. Run in the build container only:  python oracle/make_golden_ext.py  -> tests/golden/ext_cases.npz
  recon_*   src.losses.ReconstructionLoss(loss_name, use_min, use_automask, mask_name).forward on pre-warped frames
            (masks, 'l2', C-channel inputs): loss, d loss/d pred, d loss/d mask, automask
  smooth_*  src.regularizers.SmoothReg(use_edges, use_laplacian): loss, d loss/d disp, the two gradient maps
  feat_*    src.core.handlers.feat_recon with the reference's ViewSynth: loss, d loss/d depth, warped features
Inputs are regenerated in the tests from the recorded seeds (numpy RandomState, float32 values widened to float64).
`use_blur` needs kornia.filters.gaussian_blur2d (absent here): no fixture, the oracle restates it (parity unpinned)."""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
warnings.filterwarnings('ignore')

from oracle import ref_shim  # noqa: E402

GOLDEN = ROOT/'tests'/'golden'

RECON = {
    'recon_ssim_min_auto_expl': dict(loss_name='ssim', use_min=True, use_automask=True, mask_name='explainability', b=2, n=2, C=3, H=14, W=18, seed=11),
    'recon_ssim_mean_auto_unc': dict(loss_name='ssim', use_min=False, use_automask=True, mask_name='uncertainty', b=1, n=3, C=3, H=12, W=16, seed=12),
    'recon_l1_min_unc': dict(loss_name='l1', use_min=True, use_automask=False, mask_name='uncertainty', b=2, n=2, C=3, H=10, W=12, seed=13),
    'recon_l2_mean_feat': dict(loss_name='l2', use_min=False, use_automask=False, mask_name=None, b=2, n=2, C=8, H=10, W=14, seed=14),
    'recon_l2_min_auto_unc_feat': dict(loss_name='l2', use_min=True, use_automask=True, mask_name='uncertainty', b=1, n=2, C=5, H=9, W=11, seed=15),
    'recon_ssim_min_c5': dict(loss_name='ssim', use_min=True, use_automask=False, mask_name=None, b=1, n=2, C=5, H=11, W=13, seed=16),
}
SMOOTH = {
    'smooth_lap_edges': dict(use_edges=True, use_laplacian=True, b=2, H=14, W=18, seed=21),
    'smooth_lap_plain': dict(use_edges=False, use_laplacian=True, b=1, H=9, W=12, seed=22),
    'smooth_grad_plain': dict(use_edges=False, use_laplacian=False, b=2, H=8, W=10, seed=23),
}
REGR = {
    'regr_berhu_mask': dict(loss_name='berhu', invert=False, masked=True, n=(2, 1, 9, 13), seed=61),
    'regr_logl1_invert': dict(loss_name='log_l1', invert=True, masked=True, n=(2, 1, 8, 10), seed=62),
    'regr_l1_plain': dict(loss_name='l1', invert=False, masked=False, n=(1, 1, 7, 9), seed=63),
    'regr_berhu_invert_plain': dict(loss_name='berhu', invert=True, masked=False, n=(1, 1, 6, 11), seed=64),
}
STEREO = {'stereo_l1': dict(loss_name='l1', b=2, S=2, H=12, W=20, seed=71)}
HINTS = {'hints_logl1_auto': dict(loss_name='log_l1', invert=False, use_automask=True, b=2, S=2, n=2, H=16, W=24, seed=81),
         'hints_berhu_invert': dict(loss_name='berhu', invert=True, use_automask=False, b=1, S=1, n=2, H=12, W=16, seed=82)}
FREG = {
    'freg_peaky_edges': dict(cls='FeatPeakReg', use_edges=True, b=2, C=5, H=9, W=12, seed=91),
    'freg_peaky_plain': dict(cls='FeatPeakReg', use_edges=False, b=1, C=3, H=6, W=7, seed=92),
    'freg_smooth_edges': dict(cls='FeatSmoothReg', use_edges=True, b=2, C=4, H=10, W=11, seed=93),
    'freg_smooth_plain': dict(cls='FeatSmoothReg', use_edges=False, b=1, C=6, H=7, W=9, seed=94),
}
FEAT = {
    'feat_l2_mean': dict(loss_name='l2', use_min=False, use_automask=False, b=2, n=2, C=6, H=16, W=24, seed=31),
    'feat_l2_min_auto': dict(loss_name='l2', use_min=True, use_automask=True, b=1, n=2, C=4, H=16, W=20, seed=32),
}


def recon_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).to(dtype)
    tgt = f(c['b'], c['C'], c['H'], c['W'])
    pred = (tgt[None] + 0.3*(f(c['n'], c['b'], c['C'], c['H'], c['W']) - 0.5)).clamp(0, 1)
    src = (tgt[None] + 0.3*(f(c['n'], c['b'], c['C'], c['H'], c['W']) - 0.5)).clamp(0, 1)
    mask = 0.1 + 0.8*f(c['b'], c['n'], c['H'], c['W'])
    noise = torch.from_numpy(rs.standard_normal((c['b'], 1, c['H'], c['W'])).astype(np.float32)).to(dtype)
    return dict(pred=pred, tgt=tgt, src=src, mask=mask, noise=noise)


def smooth_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).to(dtype)
    return dict(disp=0.05 + 0.9*f(c['b'], 1, c['H'], c['W']), img=f(c['b'], 3, c['H'], c['W']))


def feat_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).to(dtype)
    H, W = c['H'], c['W']
    feats = f(c['b'], c['C'], H//4, W//4)
    supp = (feats[None] + 0.4*(f(c['n'], c['b'], c['C'], H//4, W//4) - 0.5))
    depth = 1.0 + 4.0*f(c['b'], 1, H, W)
    aa = torch.from_numpy((0.01*rs.standard_normal((c['n'], c['b'], 3))).astype(np.float32)).to(dtype)
    t = torch.from_numpy((0.05*rs.standard_normal((c['n'], c['b'], 3))).astype(np.float32)).to(dtype)
    K = torch.tensor([[.58*W, 0, .5*W, 0], [0, 1.92*H, .5*H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=dtype).expand(c['b'], 4, 4).clone()
    noise = torch.from_numpy(rs.standard_normal((c['b'], 1, H, W)).astype(np.float32)).to(dtype)
    return dict(feats=feats, supp=supp, depth=depth, aa=aa, t=t, K=K, noise=noise)


def regr_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    f = lambda: torch.from_numpy(rs.random_sample(c['n']).astype(np.float32)).to(dtype)
    pred, tgt = 0.2 + 3*f(), 0.2 + 3*f()
    mask = (f() > 0.3).to(dtype)
    return dict(pred=pred, tgt=tgt, mask=mask)


def stereo_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    H, W, b = c['H'], c['W'], c['b']
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).to(dtype)
    disps = [0.05 + 0.9*f(b, 1, H, W) for _ in range(c['S'])]
    disps_st = [0.05 + 0.9*f(b, 1, H, W) for _ in range(c['S'])]
    T = torch.eye(4, dtype=dtype).repeat(b, 1, 1); T[:, 0, 3] = -0.1
    K = torch.tensor([[.58*W, 0, .5*W, 0], [0, 1.92*H, .5*H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=dtype).expand(b, 4, 4).clone()
    return dict(disps=disps, disps_st=disps_st, T=T, K=K)


def hints_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    H, W, b, n = c['H'], c['W'], c['b'], c['n']
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).to(dtype)
    depths = [1.0 + 4.0*f(b, 1, H, W) for _ in range(c['S'])]
    hints = 1.0 + 4.0*f(b, 1, H, W)
    hints[:, :, :2] = 0        # invalid hints (masked out by `targets > 0`)
    imgs = f(b, 3, H, W)
    supp = (imgs[None] + 0.2*(f(n, b, 3, H, W) - 0.5)).clamp(0, 1)
    aa = torch.from_numpy((0.01*rs.standard_normal((n, b, 3))).astype(np.float32)).to(dtype)
    t = torch.from_numpy((0.05*rs.standard_normal((n, b, 3))).astype(np.float32)).to(dtype)
    K = torch.tensor([[.58*W, 0, .5*W, 0], [0, 1.92*H, .5*H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=dtype).expand(b, 4, 4).clone()
    return dict(depths=depths, hints=hints, imgs=imgs, supp=supp, aa=aa, t=t, K=K)


def freg_inputs(c: dict, dtype=torch.float64) -> dict:
    rs = np.random.RandomState(c['seed'])
    f = lambda *s: torch.from_numpy(rs.random_sample(s).astype(np.float32)).to(dtype)
    return dict(feat=2*f(c['b'], c['C'], c['H'], c['W']) - 1, img=f(c['b'], 3, c['H'], c['W']), x=0.05 + 0.9*f(c['b'], 1, c['H'], c['W']))


def main() -> None:
    ref_shim.load()
    from src import losses, regularizers
    from src.core import handlers
    from src.tools import T_from_AAt, ViewSynth
    out = {}
    orig = torch.randn_like
    for name, c in RECON.items():
        d = recon_inputs(c)
        pred, mask = d['pred'].clone().requires_grad_(), d['mask'].clone().requires_grad_()
        crit = losses.ReconstructionLoss(c['loss_name'], c['use_min'], c['use_automask'], c['mask_name'])
        torch.randn_like = lambda x: d['noise']
        try: loss, ld = crit(pred, d['tgt'], source=d['src'], mask=mask if c['mask_name'] else None)
        finally: torch.randn_like = orig
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy(); out[f'{name}/g_pred'] = pred.grad.numpy()
        if c['mask_name']: out[f'{name}/g_mask'] = mask.grad.numpy()
        if 'automask' in ld: out[f'{name}/automask'] = ld['automask'].numpy().astype(np.uint8)
        print(name, float(loss))
    for name, c in SMOOTH.items():
        d = smooth_inputs(c)
        disp = d['disp'].clone().requires_grad_()
        loss, ld = regularizers.SmoothReg(use_edges=c['use_edges'], use_laplacian=c['use_laplacian'])(disp, d['img'])
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy(); out[f'{name}/g_disp'] = disp.grad.numpy()
        out[f'{name}/disp_grad'] = ld['disp_grad'].detach().numpy(); out[f'{name}/image_grad'] = ld['image_grad'].detach().numpy()
        print(name, float(loss))
    for name, c in FEAT.items():
        d = feat_inputs(c)
        depth = d['depth'].clone().requires_grad_()
        Ts = T_from_AAt(d['aa'], d['t'])
        crit = losses.ReconstructionLoss(c['loss_name'], c['use_min'], c['use_automask'])
        torch.randn_like = lambda x: d['noise']
        try:
            loss, ld = handlers.feat_recon(crit, ViewSynth((c['H'], c['W'])).to(torch.float64), {0: depth}, None, d['feats'], d['supp'], Ts, d['K'])
        finally: torch.randn_like = orig
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy(); out[f'{name}/g_depth'] = depth.grad.numpy()
        out[f'{name}/warp'] = ld['supp_feats_warp'].detach().numpy().astype(np.float32)
        print(name, float(loss))
    from src.tools import to_scaled
    for name, c in REGR.items():
        d = regr_inputs(c)
        pred, tgt = d['pred'].clone().requires_grad_(), d['tgt'].clone().requires_grad_()
        loss, ld = losses.RegressionLoss(c['loss_name'], invert=c['invert'])(pred, tgt, d['mask'] if c['masked'] else None)
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy(); out[f'{name}/g_pred'] = pred.grad.numpy(); out[f'{name}/g_tgt'] = tgt.grad.numpy()
        out[f'{name}/err'] = ld['err_regr'].detach().numpy()
        print(name, float(loss))
    for name, c in STEREO.items():
        d = stereo_inputs(c)
        disps = {s: x.clone().requires_grad_() for s, x in enumerate(d['disps'])}
        disps_st = {s: x.clone().requires_grad_() for s, x in enumerate(d['disps_st'])}
        depths = {s: to_scaled(x, 0.1, 100.)[1] for s, x in disps.items()}
        depths_st = {s: to_scaled(x, 0.1, 100.)[1] for s, x in disps_st.items()}
        loss, ld = handlers.stereo_const(losses.RegressionLoss(c['loss_name']), ViewSynth((c['H'], c['W'])).to(torch.float64), disps, depths,
                                         disps_st, depths_st, d['T'], d['K'])
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy()
        for s in disps: out[f'{name}/g_disp{s}'] = disps[s].grad.numpy(); out[f'{name}/g_disp_st{s}'] = disps_st[s].grad.numpy()
        out[f'{name}/disps_warp'] = ld['disps_warp'].detach().numpy().astype(np.float32)
        print(name, float(loss))
    for name, c in HINTS.items():
        d = hints_inputs(c)
        depths = {s: x.clone().requires_grad_() for s, x in enumerate(d['depths'])}
        photo = losses.ReconstructionLoss('ssim', use_min=True).compute_photo
        crit = losses.RegressionLoss(c['loss_name'], invert=c['invert'], use_automask=c['use_automask'])
        loss, ld = handlers.depth_regr(crit, ViewSynth((c['H'], c['W'])).to(torch.float64), photo, depths, d['hints'], d['imgs'], d['supp'],
                                       T_from_AAt(d['aa'], d['t']), d['K'])
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy()
        for s in depths: out[f'{name}/g_depth{s}'] = depths[s].grad.numpy()
        out[f'{name}/mask'] = ld['mask_regr'].numpy().astype(np.uint8)
        print(name, float(loss), float(ld['mask_regr'].float().mean()))
    for name, c in FREG.items():
        d = freg_inputs(c)
        feat = d['feat'].clone().requires_grad_()
        loss, ld = getattr(regularizers, c['cls'])(use_edges=c['use_edges'])(feat, d['img'])
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy(); out[f'{name}/g_feat'] = feat.grad.numpy()
        out[f'{name}/feat_grad'] = ld['feat_grad'].detach().numpy()
        print(name, float(loss))
    d = freg_inputs(FREG['freg_peaky_edges'])
    for name, crit in (('pw_mask', regularizers.MaskReg()), ('pw_occ', regularizers.OccReg()), ('pw_occ_inv', regularizers.OccReg(invert=True))):
        x = d['x'].clone().requires_grad_()
        loss, _ = crit(x)
        loss.backward()
        out[f'{name}/loss'] = loss.detach().numpy(); out[f'{name}/g_x'] = x.grad.numpy()
        print(name, float(loss))
    np.savez_compressed(GOLDEN/'ext_cases.npz', **out)
    print('->', GOLDEN/'ext_cases.npz', (GOLDEN/'ext_cases.npz').stat().st_size//1024, 'KiB')


if __name__ == '__main__':
    main()
