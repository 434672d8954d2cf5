"""CPU: the drop-in boundary. With the REAL reference importable (this build container; skipped on the GPU box, where
/root/reference does not exist) `plugin.install()` must put the B200 classes behind the reference's own registry keys and
call-time names, the reference's `MonoDepthModule(cfg)` must then construct itself from them unchanged, and `uninstall()` must
restore every original."""
import warnings

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')

CFG = {
    'net': {'depth': {'enc_name': 'convnext_tiny', 'pretrained': False, 'dec_name': 'monodepth', 'out_scales': [0, 1, 2, 3]},
            'pose': {'enc_name': 'resnet18', 'pretrained': False, 'learn_K': True}},
    'loss': {'img_recon': {'weight': 1, 'loss_name': 'ssim', 'use_min': True, 'use_automask': True},
             'disp_smooth': {'weight': 0.001, 'use_edges': True}},
    'optimizer': {'type': 'adamw', 'lr': 1e-4}, 'scheduler': None, 'dataset': {}, 'loader': {'batch_size': 8},
    'trainer': {'min_depth': 0.1, 'max_depth': 100, 'aspect_ratio_aug_prob': 0.7, 'aspect_ratio_ref_shape': (384, 640)},
}


def test_install_rebinds_the_reference_hooks_and_uninstall_restores_them():
    warnings.filterwarnings('ignore')
    ref_shim.load()
    import src.core.handlers as rh
    import src.core.trainer as rt
    import src.registry as reg
    from slowtv_monodepth_b200 import aspect_ratio, geometry, handlers, losses, networks, plugin, regularizers
    reg.trigger_nets(); reg.trigger_decoders(); reg.trigger_losses()
    before = {'depth': reg.NET_REG['depth'], 'pose': reg.NET_REG['pose'], 'monodepth': reg.DEC_REG['monodepth'],
              'img_recon': reg.LOSS_REG['img_recon'], 'disp_smooth': reg.LOSS_REG['disp_smooth'], 'ViewSynth': rt.ViewSynth,
              'aspect_ratio_aug': rt.aspect_ratio_aug, 'image_recon': rh.image_recon, 'disp_smooth_fn': rh.disp_smooth}
    assert before['depth'].__module__.startswith('src.')
    plugin.install()
    try:
        assert reg.NET_REG['depth'] is networks.DepthNet and reg.NET_REG['pose'] is networks.PoseNet
        assert reg.DEC_REG['monodepth'] is networks.MonodepthDecoder
        assert reg.LOSS_REG['img_recon'] is losses.ReconstructionLoss and reg.LOSS_REG['disp_smooth'] is regularizers.SmoothReg
        assert reg.LOSS_REG['feat_recon'] is losses.ReconstructionLoss and reg.LOSS_REG['autoenc_recon'] is losses.ReconstructionLoss  # one class, three keys
        assert reg.LOSS_REG['depth_regr'] is losses.RegressionLoss and reg.LOSS_REG['stereo_const'] is losses.RegressionLoss
        assert rh.feat_recon is handlers.feat_recon and rh.stereo_const is handlers.stereo_const and rh.depth_regr is handlers.depth_regr
        assert reg.LOSS_REG['feat_peaky'] is regularizers.FeatPeakReg and reg.LOSS_REG['feat_smooth'] is regularizers.FeatSmoothReg
        assert reg.LOSS_REG['disp_mask'] is regularizers.MaskReg and reg.LOSS_REG['disp_occ'] is regularizers.OccReg
        assert rt.ViewSynth is geometry.ViewSynth and rt.aspect_ratio_aug is aspect_ratio.aspect_ratio_aug
        assert rh.image_recon is handlers.image_recon and rh.disp_smooth is handlers.disp_smooth

        # the reference's own module builds itself from the replaced pieces, through its own parsers (src/tools/parsers.py:68,103)
        module = rt.MonoDepthModule(CFG)
        assert isinstance(module.nets['depth'], networks.DepthNet) and isinstance(module.nets['pose'], networks.PoseNet)
        assert isinstance(module.losses['img_recon'], losses.ReconstructionLoss) and module.weights['disp_smooth'] == 0.001
        assert module.ar_aug.func is aspect_ratio.aspect_ratio_aug and module.ar_aug.keywords['p'] == 0.7
        assert module.scales == [0, 1, 2, 3]
        # parameter names follow the reference's state_dict layout: a reference checkpoint's keys resolve
        ref_names = set(before['depth'](**CFG['net']['depth']).state_dict())
        assert ref_names == set(module.nets['depth'].state_dict())
    finally:
        plugin.uninstall()
    assert reg.NET_REG['depth'] is before['depth'] and reg.DEC_REG['monodepth'] is before['monodepth']
    assert reg.LOSS_REG['img_recon'] is before['img_recon'] and rt.ViewSynth is before['ViewSynth']
    assert rt.aspect_ratio_aug is before['aspect_ratio_aug'] and rh.image_recon is before['image_recon'] and rh.disp_smooth is before['disp_smooth_fn']
