"""CPU: the oracle's DepthNet / MonodepthDecoder / PoseNet restatements equal the REFERENCE's own classes (run on top of the
oracle's timm stand-in through oracle/ref_shim.py) for shared weights, and the product networks equal the oracle's.
The reference-backed part needs /root/reference and is skipped where that tree is absent (e.g. on the GPU box)."""
import warnings

import pytest
import torch

from oracle import nets as ON, ref_shim

warnings.filterwarnings('ignore')


def _x(b, c, h, w, seed=0):
    return torch.randn(b, c, h, w, generator=torch.Generator().manual_seed(seed))


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')
@pytest.mark.parametrize('enc', ['resnet18', 'convnext_tiny'])
def test_oracle_nets_equal_reference_classes(enc):
    ref_shim.load()
    from src.registry import NET_REG, trigger_decoders, trigger_nets
    trigger_nets(); trigger_decoders()
    torch.manual_seed(0)
    rd = NET_REG['depth'](enc_name=enc, pretrained=False, dec_name='monodepth', out_scales=[0, 1, 2, 3])
    rp = NET_REG['pose'](enc_name='resnet18', learn_K=True, pretrained=False)
    od, op = ON.DepthNet(enc), ON.PoseNet('resnet18', learn_K=True)
    assert set(od.state_dict()) == set(rd.state_dict()) and set(op.state_dict()) == set(rp.state_dict())
    od.load_state_dict(rd.state_dict()); op.load_state_dict(rp.state_dict())
    x = _x(2, 3, 64, 96)
    a, b = rd(x), od(x)
    for s in range(4): assert torch.allclose(a['disp'][s], b['disp'][s], atol=1e-6), s
    x6 = _x(2, 6, 64, 96, 1)
    a, b = rp(x6), op(x6)
    for k in ('R', 't', 'fs', 'cs'): assert torch.allclose(a[k], b[k], atol=1e-7), k


@pytest.mark.parametrize('enc', ['resnet18', 'convnext_tiny'])
def test_product_nets_equal_oracle_nets_on_cpu(enc):
    """Same parameter names, same arithmetic (host tensors; the GPU variant lives in test_nets_gpu.py)."""
    from slowtv_monodepth_b200.networks import DepthNet, PoseNet
    torch.manual_seed(1)
    od, op = ON.DepthNet(enc).double(), ON.PoseNet('resnet18', learn_K=True).double()
    pd, pp = DepthNet(enc, pretrained=False).double(), PoseNet('resnet18', learn_K=True).double()
    assert list(pd.state_dict()) and set(pd.state_dict()) == set(od.state_dict())
    assert set(pp.state_dict()) == set(op.state_dict())
    pd.load_state_dict(od.state_dict()); pp.load_state_dict(op.state_dict())
    x = _x(2, 3, 64, 96).double().requires_grad_()
    a, b = od(x), pd(x)
    for s in range(4): assert torch.allclose(a['disp'][s], b['disp'][s], atol=1e-10), s
    ga = torch.autograd.grad(sum(v.sum() for v in a['disp'].values()), list(od.parameters()))
    gb = torch.autograd.grad(sum(v.sum() for v in b['disp'].values()), list(pd.parameters()))
    names = [n for n, _ in od.named_parameters()]
    pn = dict(zip([n for n, _ in pd.named_parameters()], gb))
    for n, g in zip(names, ga): assert torch.allclose(g, pn[n], atol=1e-8, rtol=1e-6), n
    x6 = _x(2, 6, 64, 96, 1).double()
    a, b = op(x6), pp(x6)
    for k in ('R', 't', 'fs', 'cs'): assert torch.allclose(a[k], b[k], atol=1e-12), k
