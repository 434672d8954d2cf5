"""CPU: checkpoint compatibility. With the real reference importable, its own `MonoDepthModule` (original
classes) writes a Lightning-style checkpoint; the B200 modules must load it strictly (same names, same shapes, same values), the
quickstart path must rebuild a DepthNet from the stored hyper-parameters, and a checkpoint written by `save_nets` must load back
into the REFERENCE's classes. Skipped where /root/reference does not exist (GPU box)."""
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')

CFG = {
    'net': {'depth': {'enc_name': 'convnext_tiny', 'pretrained': False, 'dec_name': 'monodepth', 'out_scales': [0, 1, 2, 3]},
            'pose': {'enc_name': 'resnet18', 'pretrained': False, 'learn_K': True}},
    'loss': {'img_recon': {'weight': 1, 'loss_name': 'ssim', 'use_min': True, 'use_automask': True},
             'disp_smooth': {'weight': 0.001, 'use_edges': True}},
    'optimizer': {'type': 'adamw', 'lr': 1e-4}, 'scheduler': None, 'dataset': {}, 'loader': {'batch_size': 8},
    'trainer': {'min_depth': 0.1, 'max_depth': 100},
}


def test_reference_checkpoint_round_trip(tmp_path):
    warnings.filterwarnings('ignore')
    ref_shim.load()
    import src.core.trainer as rt
    from slowtv_monodepth_b200 import checkpoint as CK
    from slowtv_monodepth_b200.trainer import MonoDepthStep, default_cfg
    torch.manual_seed(3)
    ref = rt.MonoDepthModule(CFG)                                   # the reference's own networks
    assert type(ref.nets['depth']).__module__.startswith('src.')
    with torch.no_grad():
        for p in ref.parameters():
            if p.is_floating_point(): p.add_(0.01*torch.randn_like(p))    # not the init values
    path = tmp_path/'ref.ckpt'
    torch.save({'state_dict': ref.state_dict(), 'hyper_parameters': {'cfg': CFG}}, path)

    ours = MonoDepthStep(default_cfg('convnext_tiny', 'resnet18', learn_K=True))
    assert CK.load_nets(ours.nets, path) == ['depth', 'pose']
    want = ref.nets.state_dict()
    got = ours.nets.state_dict()
    assert set(got) == set(want)
    for k, v in want.items(): assert torch.equal(got[k].cpu(), v), k

    depth = CK.load_depth_net(path)                                  # api/quickstart/run.py:21-33
    assert not any(p.requires_grad for p in depth.parameters()) and not depth.training
    for k, v in ref.nets['depth'].state_dict().items(): assert torch.equal(depth.state_dict()[k], v), k

    back = tmp_path/'ours.ckpt'
    CK.save_nets(ours.nets, back, cfg=CFG)
    ck = torch.load(back, map_location='cpu', weights_only=False)
    fresh = rt.MonoDepthModule(CFG)
    missing, unexpected = fresh.load_state_dict(ck['state_dict'], strict=False)
    assert not unexpected and all(not m.startswith('nets.') for m in missing)      # the reference reads every network weight back
    for k, v in ref.nets.state_dict().items(): assert torch.equal(fresh.nets.state_dict()[k], v), k


def test_errors():
    from slowtv_monodepth_b200 import checkpoint as CK
    from slowtv_monodepth_b200.trainer import MonoDepthStep, default_cfg
    with pytest.raises(ValueError): CK.load_nets(torch.nn.ModuleDict(), {'weights': {}})
    ours = MonoDepthStep(default_cfg('resnet18', 'resnet18'))
    with pytest.raises(KeyError): CK.load_nets(ours.nets, {'state_dict': {'nets.autoencoder.x': torch.zeros(1)}, 'hyper_parameters': {}})
    with pytest.raises(RuntimeError): CK.load_nets(ours.nets, {'state_dict': {'nets.depth.x': torch.zeros(1), 'nets.pose.y': torch.zeros(1)}})
    assert CK.split_state_dict({'nets.depth.a.b': 1, 'losses.x': 2, 'nets.pose.c': 3}) == {'depth': {'a.b': 1}, 'pose': {'c': 3}}
