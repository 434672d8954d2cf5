"""GPU: tcgen05 TF32 GEMM (stv_gemm_tf32, through the C ABI) vs float64 matmul.

Two kinds of check:
  * layout exactness — operands hold small integers (exactly representable in TF32, products and sums exact in fp32), so
    ANY mistake in the TMA boxes / swizzle / UMMA descriptors / TMEM read-back shows up as a non-zero difference;
  * TF32 accuracy — random fp32 operands: |C - C64| <= 2^-10 * sum_k |a||b| (TF32 keeps 10 mantissa bits per operand;
    the hardware truncates, giving a relative error < 2^-10 per factor, < 2^-9 per product).
"""
import pytest
import torch

from slowtv_monodepth_b200 import functional as F_

pytestmark = pytest.mark.gpu

MAJORS = [(False, False), (False, True), (True, False), (True, True)]


def _ints(shape, gen):
    return torch.randint(-3, 4, shape, generator=gen, device='cuda').float()


def _mn_storage(X):
    """(rows, K) logical operand -> its MN-major storage (K, rows) with the row pitch padded to a multiple of 4 floats (TMA)."""
    rows, K = X.shape
    buf = torch.zeros(K, (rows + 3)//4*4, device=X.device)
    buf[:, :rows] = X.t()
    return buf[:, :rows]


def _operands(M, N, K, a_mn, b_mn, gen, ints):
    mk = (lambda s: _ints(s, gen)) if ints else (lambda s: torch.randn(s, generator=gen, device='cuda'))
    A, B = mk((M, K)), mk((N, K))
    return A, B, (_mn_storage(A) if a_mn else A), (_mn_storage(B) if b_mn else B)


@pytest.mark.parametrize('a_mn,b_mn', MAJORS)
@pytest.mark.parametrize('M,N,K', [(128, 32, 32), (256, 128, 64), (200, 96, 80), (1000, 384, 96), (77, 768, 520), (513, 260, 36)])
def test_exact_on_integers(M, N, K, a_mn, b_mn):
    gen = torch.Generator(device='cuda').manual_seed(M*7 + N*3 + K)
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn, gen, ints=True)
    got = F_.gemm_tf32(As, Bs, a_mn=a_mn, b_mn=b_mn)
    want = (A.double() @ B.double().t()).float()
    assert torch.equal(got, want), f'max |diff| = {(got - want).abs().max().item()}'


@pytest.mark.parametrize('a_mn,b_mn', MAJORS)
def test_tf32_accuracy(a_mn, b_mn):
    gen = torch.Generator(device='cuda').manual_seed(5)
    M, N, K = 640, 192, 1024
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn, gen, ints=False)
    got = F_.gemm_tf32(As, Bs, a_mn=a_mn, b_mn=b_mn).double()
    want = A.double() @ B.double().t()
    bound = (A.abs().double() @ B.abs().double().t())*2.0**-9
    assert ((got - want).abs() <= bound + 1e-6).all()
    assert (got - want).norm()/want.norm() < 1e-3


def test_strided_operands_and_output():
    gen = torch.Generator(device='cuda').manual_seed(9)
    M, N, K = 300, 64, 96
    bigA, bigB = _ints((M, K + 32), gen), _ints((N, K + 8), gen)
    out = torch.zeros(M, N + 16, device='cuda')
    F_.gemm_tf32(bigA[:, :K], bigB[:, :K], out=out[:, :N])
    want = (bigA[:, :K].double() @ bigB[:, :K].double().t()).float()
    assert torch.equal(out[:, :N], want) and (out[:, N:] == 0).all()


@pytest.mark.parametrize('act', ['none', 'relu', 'gelu', 'elu', 'sigmoid'])
def test_forward_epilogue(act):
    gen = torch.Generator(device='cuda').manual_seed(11)
    M, N, K = 333, 160, 64
    A, B = _ints((M, K), gen)*0.25, _ints((N, K), gen)*0.25
    bias, gamma, res = torch.randn(N, generator=gen, device='cuda'), torch.randn(N, generator=gen, device='cuda'), torch.randn(M, N, generator=gen, device='cuda')
    aux = torch.empty(M, N, device='cuda')
    got = F_.gemm_tf32(A, B, bias=bias, act=act, aux=aux, gamma=gamma, res=res)
    z = (A.double() @ B.double().t()) + bias.double()
    fn = {'none': lambda x: x, 'relu': torch.relu, 'gelu': torch.nn.functional.gelu, 'elu': torch.nn.functional.elu, 'sigmoid': torch.sigmoid}[act]
    want = fn(z)*gamma.double() + res.double()
    assert (aux.double() - z).abs().max() < 1e-5
    assert (got.double() - want).abs().max() < 2e-5


@pytest.mark.parametrize('act', ['relu', 'gelu', 'elu', 'sigmoid'])
def test_activation_backward_epilogue(act):
    gen = torch.Generator(device='cuda').manual_seed(13)
    M, N, K = 260, 96, 128
    A, B = _ints((M, K), gen)*0.25, _ints((N, K), gen)*0.25
    pre = torch.randn(M, N, generator=gen, device='cuda', dtype=torch.float64).requires_grad_()
    fn = {'relu': torch.relu, 'gelu': torch.nn.functional.gelu, 'elu': torch.nn.functional.elu, 'sigmoid': torch.sigmoid}[act]
    y = fn(pre)
    g = A.double() @ B.double().t()
    want, = torch.autograd.grad(y, pre, g)
    src = (pre if act == 'gelu' else y).detach().float()
    got = F_.gemm_tf32(A, B, dact=act, dact_src=src)
    assert (got.double() - want).abs().max() < 1e-4*max(1.0, want.abs().max().item())


@pytest.mark.parametrize('split_k', [1, 3, 16])
def test_split_k_accumulate(split_k):
    gen = torch.Generator(device='cuda').manual_seed(17)
    M, N, K = 96, 384, 4000  # weight-gradient shape: small output, long reduction
    A, B, As, Bs = _operands(M, N, K, True, True, gen, ints=True)
    out = torch.ones(M, N, device='cuda')
    F_.gemm_tf32(As, Bs, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=split_k)
    want = (A.double() @ B.double().t()).float() + 1
    assert torch.equal(out, want)


def test_colsum_epilogue():
    gen = torch.Generator(device='cuda').manual_seed(29)
    M, N, K = 333, 200, 64
    A, B = _ints((M, K), gen), _ints((N, K), gen)
    cs = torch.ones(N, device='cuda')
    got = F_.gemm_tf32(A, B, colsum=cs)
    want = (A.double() @ B.double().t())
    assert torch.equal(got, want.float()) and torch.equal(cs, (want.sum(0) + 1).float())


def test_bad_arguments_raise():
    A, B = torch.zeros(8, 6, device='cuda'), torch.zeros(8, 6, device='cuda')
    with pytest.raises(ValueError): F_.gemm_tf32(A, B)  # K = 6 -> lda not a multiple of 4
    with pytest.raises(ValueError): F_.gemm_tf32(torch.zeros(8, 8, device='cuda'), torch.zeros(8, 12, device='cuda'))


@pytest.mark.parametrize('act', [None, 'relu'])
def test_linear_autograd(act):
    gen = torch.Generator(device='cuda').manual_seed(23)
    M, K, N = 520, 48, 96
    x, w, b = _ints((M, K), gen).requires_grad_(), _ints((N, K), gen).requires_grad_(), _ints((N,), gen).requires_grad_()
    y = F_.linear(x, w, b, act)
    dy = _ints((M, N), gen)
    y.backward(dy)
    xr, wr, br = (t.detach().double().requires_grad_() for t in (x, w, b))
    yr = xr @ wr.t() + br
    if act == 'relu': yr = torch.relu(yr)
    yr.backward(dy.double())
    assert torch.equal(y, yr.float())
    for a, r in ((x, xr), (w, wr), (b, br)): assert torch.equal(a.grad, r.grad.float())
