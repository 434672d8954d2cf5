"""GPU: tcgen05 implicit-GEMM convolutions (stv_conv_fprop / dgrad / wgrad + stv_grad_pull / stv_act_bwd, through the C ABI) vs
torch's float64 conv2d on the explicitly materialised virtual input (nearest x2 upsample -> concat -> reflect / zero pad).

Operands are small integers (or quarter-integers), exactly representable in TF32 with exact fp32 sums, so the linear cases
must match BIT-EXACTLY: any error in the gather indexing, swizzles, descriptors, padding / fold / pool logic is visible.
"""
import pytest
import torch
import torch.nn.functional as F

from slowtv_monodepth_b200 import functional as F_

pytestmark = pytest.mark.gpu

#        N  H   W   C1  C2  up1    Cout R  st pad reflect
CASES = {
    'zero3x3':        (2, 12, 20, 32, 0, False, 64, 3, 1, 1, False),
    'reflect3x3_c16': (2, 10, 14, 16, 0, False, 16, 3, 1, 1, True),
    'up_cat_reflect': (2, 12, 16, 32, 64, True, 32, 3, 1, 1, True),
    'up_reflect':     (1, 8, 12, 32, 0, True, 16, 3, 1, 1, True),
    'stride2_3x3':    (2, 12, 16, 64, 0, False, 128, 3, 2, 1, False),
    'stride2_1x1':    (2, 12, 16, 64, 0, False, 128, 1, 2, 0, False),
    'stem7x7_s2':     (2, 20, 28, 8, 0, False, 64, 7, 2, 3, False),
    'patch4x4_s4':    (2, 16, 24, 4, 0, False, 96, 4, 4, 0, False),
    'down2x2_s2':     (2, 12, 16, 96, 0, False, 192, 2, 2, 0, False),
    'odd_sizes':      (3, 9, 11, 48, 0, False, 40, 3, 1, 1, True),
    'wide_multi_tile': (2, 40, 56, 64, 0, False, 288, 3, 1, 1, False),
    'head_cout1':     (2, 10, 14, 16, 0, False, 1, 3, 1, 1, True),
    'head_cout1_c32': (2, 10, 14, 32, 0, False, 1, 3, 1, 1, True),
    'cat16_16_up':    (2, 12, 16, 16, 16, True, 32, 3, 1, 1, True),
    'zero3x3_c16':    (2, 12, 20, 16, 0, False, 32, 3, 1, 1, False),
    'pose_head_1x1':  (3, 6, 10, 256, 0, False, 12, 1, 1, 0, False),
    'tall_batch':     (5, 7, 33, 64, 0, False, 64, 3, 1, 1, False),
    'reflect_cout16': (2, 16, 24, 32, 0, False, 16, 3, 1, 1, True),
    # full-size rows (also re-run through the row-segment kernel below): full segments, and a partly empty last segment behind a
    # materialised reflection pad
    'rowseg_256':     (2, 200, 256, 32, 0, False, 32, 3, 1, 1, False),
    'rowseg_reflect': (1, 260, 386, 32, 0, False, 16, 3, 1, 1, True),
}

# Cases re-run with the row-segment kernel forced wherever its geometry allows (3 taps along x, stride 1, 32-channel blocks).
ROWSEG_FORCED = ['zero3x3', 'up_cat_reflect', 'reflect3x3_c16', 'tall_batch', 'reflect_cout16', 'odd_sizes', 'zero3x3_c16', 'rowseg_256', 'rowseg_reflect']
ROWSEG_EXTRA = {
    'seg_300':       (2, 6, 300, 32, 0, False, 64, 3, 1, 1, False),    # 128 + 128 + 44 pixels
    'seg_reflect':   (1, 5, 162, 32, 0, False, 32, 3, 1, 1, True),     # padded rows of 164
    'seg_cout96':    (1, 4, 140, 64, 0, False, 96, 3, 1, 1, False),
    'seg_cin64':     (2, 7, 130, 64, 0, False, 64, 3, 1, 1, False),    # two channel blocks, two column tiles, partial row block (7 = 4 + 3)
    'seg_tall':      (1, 21, 40, 32, 0, False, 16, 3, 1, 1, True),     # many row blocks per CTA: accumulator double-buffering wraps
}
CASES_ALL = {**CASES, **ROWSEG_EXTRA}


@pytest.fixture
def force_rowseg(monkeypatch):
    monkeypatch.setenv('STV_CONV_ROWSEG', '2')
    yield



def _ints(shape, gen, lo=-2, hi=3):
    return torch.randint(lo, hi, shape, generator=gen, device='cuda').float()


def _make(case, gen, scale=1.0):
    N, H, W, C1, C2, up1, Cout, R, st, pad, refl = CASES_ALL[case]
    s1 = _ints((N, H//2, W//2, C1) if up1 else (N, H, W, C1), gen)*scale
    s2 = _ints((N, H, W, C2), gen)*scale if C2 else None
    w = (_ints((Cout, C1 + C2, R, R), gen)*scale).contiguous(memory_format=torch.channels_last)
    b = _ints((Cout,), gen)*scale
    return s1, s2, w, b, dict(up1=up1, stride=st, pad=pad, reflect=refl)


def _ref(s1, s2, w, b, up1, stride, pad, reflect, act=None):
    x = s1.permute(0, 3, 1, 2).double()
    if up1: x = F.interpolate(x, scale_factor=2, mode='nearest')
    if s2 is not None: x = torch.cat([x, s2.permute(0, 3, 1, 2).double()], 1)
    if reflect and pad: x, pad = F.pad(x, (pad,)*4, mode='reflect'), 0
    y = F.conv2d(x, w.double(), b.double() if b is not None else None, stride, pad)
    y = {None: lambda v: v, 'relu': torch.relu, 'elu': F.elu, 'sigmoid': torch.sigmoid}[act](y)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize('case', list(CASES))
def test_fprop_exact(case):
    gen = torch.Generator(device='cuda').manual_seed(sum(map(ord, case)))
    s1, s2, w, b, kw = _make(case, gen)
    got = F_.conv2d_nhwc(s1, w, b, src2=s2, **kw)
    want = _ref(s1, s2, w, b, **kw).float()
    assert got.shape == want.shape
    assert torch.equal(got, want), f'max |diff| = {(got - want).abs().max().item()}'


@pytest.mark.parametrize('case', list(CASES))
def test_backward_exact(case):
    gen = torch.Generator(device='cuda').manual_seed(sum(map(ord, case)) + 1)
    s1, s2, w, b, kw = _make(case, gen)
    leaves = [t.clone().requires_grad_() for t in (s1, s2, w, b) if t is not None]
    a1, a2, aw, ab = (leaves[0], leaves[1], leaves[2], leaves[3]) if s2 is not None else (leaves[0], None, leaves[1], leaves[2])
    y = F_.conv2d_nhwc(a1, aw, ab, src2=a2, **kw)
    dA = _ints(tuple(y.shape), gen, -1, 2)
    y.backward(dA)
    refs = [t.detach().double().requires_grad_() for t in (s1, s2, w, b) if t is not None]
    r1, r2, rw, rb = (refs[0], refs[1], refs[2], refs[3]) if s2 is not None else (refs[0], None, refs[1], refs[2])
    _ref(r1, r2, rw, rb, **kw).backward(dA.double())
    for name, g, r in (('d_src1', a1, r1), ('d_src2', a2, r2), ('d_w', aw, rw), ('d_b', ab, rb)):
        if g is None: continue
        assert g.grad is not None, name
        assert torch.equal(g.grad, r.grad.float()), f'{name}: max |diff| = {(g.grad - r.grad.float()).abs().max().item()}'


@pytest.mark.parametrize('case', ROWSEG_FORCED + list(ROWSEG_EXTRA))
def test_rowseg_kernel_exact(case, force_rowseg):
    """Forward and backward through the row-segment kernel (forced): bit-exact like the im2col path."""
    test_fprop_exact(case)
    test_backward_exact(case)


@pytest.mark.parametrize('act', ['relu', 'elu', 'sigmoid'])
def test_activation_forward_backward(act):
    gen = torch.Generator(device='cuda').manual_seed(3)
    case = 'head_cout1' if act == 'sigmoid' else 'up_cat_reflect'
    s1, s2, w, b, kw = _make(case, gen, scale=0.125)
    leaves = [t.clone().requires_grad_() for t in (s1, s2, w, b) if t is not None]
    a1, a2, aw, ab = (leaves[0], leaves[1], leaves[2], leaves[3]) if s2 is not None else (leaves[0], None, leaves[1], leaves[2])
    y = F_.conv2d_nhwc(a1, aw, ab, src2=a2, act=act, **kw)
    dA = _ints(tuple(y.shape), gen, -1, 2)
    y.backward(dA)
    refs = [t.detach().double().requires_grad_() for t in (s1, s2, w, b) if t is not None]
    r1, r2, rw, rb = (refs[0], refs[1], refs[2], refs[3]) if s2 is not None else (refs[0], None, refs[1], refs[2])
    yr = _ref(r1, r2, rw, rb, act=act, **kw)
    yr.backward(dA.double())
    assert (y.double() - yr).abs().max() < 1e-5
    # The backward products run in TF32: dZ = dA*act'(y) is no longer TF32-exact, so compare norm-wise at TF32 accuracy.
    for name, g, r in (('d_src1', a1, r1), ('d_src2', a2, r2), ('d_w', aw, rw), ('d_b', ab, rb)):
        if g is None: continue
        err = (g.grad.double() - r.grad).norm()/r.grad.norm().clamp(min=1e-12)
        assert err < 2e-3, f'{name}: rel err {err.item():.3e}'


@pytest.mark.parametrize('N,H,W,C', [(2, 10, 14, 16), (1, 7, 9, 32), (3, 6, 8, 64), (2, 5, 6, 128), (1, 33, 65, 16), (1, 3, 3, 16), (2, 4, 5, 8),
                                     (1, 2, 2, 16), (2, 64, 300, 16), (1, 24, 37, 128)])
def test_head3x3_matches_conv2d(N, H, W, C):
    """stv_head3x3_fwd/bwd (one-channel reflect-padded 3x3 conv + sigmoid on the CUDA cores) vs torch in float64."""
    gen = torch.Generator(device='cuda').manual_seed(C + H)
    x = (torch.randn(N, H, W, C, generator=gen, device='cuda')).requires_grad_()
    w = (torch.randn(1, C, 3, 3, generator=gen, device='cuda')*0.2).requires_grad_()
    b = torch.randn(1, generator=gen, device='cuda').requires_grad_()
    dA = torch.randn(N, H, W, 1, generator=gen, device='cuda')
    y = F_.head3x3(x, w, b, 'sigmoid')
    y.backward(dA)
    xr, wr, br = (t.detach().double().requires_grad_() for t in (x, w, b))
    yr = torch.sigmoid(F.conv2d(F.pad(xr.permute(0, 3, 1, 2), (1, 1, 1, 1), mode='reflect'), wr, br)).permute(0, 2, 3, 1)
    yr.backward(dA.double())
    assert (y.double() - yr).abs().max() < 1e-5
    for name, a, r in (('dx', x, xr), ('dw', w, wr), ('db', b, br)):
        err = (a.grad.double() - r.grad).norm()/r.grad.norm()
        assert err < 5e-5, f"{name}: {err.item():.3e}"  # fp32 atomics: summation order varies
