"""CPU: the host-side algebra of SURVEY 8a row 5 (PyTorch autograd on (b,4,4) matrices: `T_from_AAt`, `to_scaled` / `to_inv`,
`resize_K`, `centre_crop_K`, `PoseNet.build_K`) against the oracle restatement — and, where the reference tree is importable
(this build container), against the reference's own functions (src/tools/geometry.py:62-90,181-263, src/networks/pose.py:61-73)
— values and gradients in float64."""
import warnings

import pytest
import torch

from oracle import loss as OL, ref_shim
from slowtv_monodepth_b200 import aspect_ratio as AR, geometry as G
from slowtv_monodepth_b200.networks import PoseNet


def _inputs():
    g = torch.Generator().manual_seed(4)
    aa = torch.randn(2, 3, 3, generator=g, dtype=torch.float64)*0.3
    aa[0, 0] = 0.                                        # zero rotation: the eps-clamped axis (geometry.py:136-140)
    t = torch.randn(2, 3, 3, generator=g, dtype=torch.float64)
    K = torch.eye(4, dtype=torch.float64).repeat(3, 1, 1)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = 0.58*640, 1.92*384, 0.5*640 + 3, 0.5*384 - 2
    return aa, t, K


def _grad(fn, *xs):
    xs = [x.clone().requires_grad_() for x in xs]
    out = fn(*xs)
    (out*torch.linspace(0.5, 1.5, out.numel(), dtype=out.dtype).view(out.shape)).sum().backward()
    return out.detach(), [x.grad for x in xs]


def test_against_the_oracle():
    aa, t, K = _inputs()
    a, ga = _grad(G.T_from_AAt, aa, t)
    b, gb = _grad(OL.T_from_AAt, aa, t)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-14) and all(torch.allclose(x, y, rtol=1e-10, atol=1e-12) for x, y in zip(ga, gb))
    assert torch.allclose(G.resize_K(K, (192, 320), (384, 640)), OL.resize_K(K, (192, 320), (384, 640)))
    fs, cs = torch.rand(5, 2, dtype=torch.float64) + 0.5, torch.rand(5, 2, dtype=torch.float64)
    assert torch.equal(PoseNet.build_K(fs, cs), OL.build_K(fs, cs))
    disp = torch.rand(2, 1, 5, 7, dtype=torch.float64)
    assert torch.allclose(G.to_scaled(disp, 0.1, 100.)[1], OL.disp_to_depth(disp, 0.1, 100.), rtol=1e-12)
    with pytest.raises(ValueError): G.T_from_AAt(aa[..., :2], t)
    with pytest.raises(ValueError): G.to_scaled(disp, -1.0, 100.)


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')
def test_against_the_reference():
    warnings.filterwarnings('ignore')
    ref_shim.load()
    from src.networks.pose import PoseNet as RefPose
    from src.tools import geometry as RG
    aa, t, K = _inputs()
    a, ga = _grad(G.T_from_AAt, aa, t)
    b, gb = _grad(RG.T_from_AAt, aa, t)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-14) and all(torch.allclose(x, y, rtol=1e-10, atol=1e-12) for x, y in zip(ga, gb))
    assert torch.allclose(G.resize_K(K, (192, 320), (384, 640)), RG.resize_K(K, (192, 320), shape=(384, 640)))
    assert torch.allclose(G.resize_K(K, (192, 320)), RG.resize_K(K, (192, 320)))
    assert torch.allclose(AR.centre_crop_K(K, (245, 588), (384, 640)), RG.centre_crop_K(K, (245, 588), (384, 640)))
    disp = torch.rand(2, 1, 5, 7, dtype=torch.float64)
    for ours, theirs in zip(G.to_scaled(disp, 0.1, 100.), RG.to_scaled(disp, 0.1, 100.)): assert torch.allclose(ours, theirs, rtol=1e-12)
    assert torch.allclose(G.to_inv(disp - 0.5), RG.to_inv(disp - 0.5))
    fs, cs = torch.rand(5, 2) + 0.5, torch.rand(5, 2)   # float32: the reference builds K from a float32 identity (pose.py:68)
    assert torch.equal(PoseNet.build_K(fs, cs), RefPose.build_K(fs, cs))
