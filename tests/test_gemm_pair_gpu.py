"""GPU: the CTA-pair (cta_group::2, 256-row tiles) variant of the tcgen05 GEMM. The library picks it for plain matrices with enough
256-row tiles to fill the SMs; these shapes are large enough to take that path (tests/test_gemm_gpu.py covers the one-CTA kernel)."""
import pytest
import torch

from slowtv_monodepth_b200 import functional as F_
from tests.test_gemm_gpu import MAJORS, _ints, _operands

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('a_mn,b_mn', MAJORS)
@pytest.mark.parametrize('M,N,K', [(7680, 1536, 384), (7680, 384, 1536), (30720, 768, 192), (9999, 640, 72), (4100, 1100, 200)])
def test_pair_exact_on_integers(M, N, K, a_mn, b_mn):
    gen = torch.Generator(device='cuda').manual_seed(M + N*3 + K)
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn, gen, ints=True)
    got = F_.gemm_tf32(As, Bs, a_mn=a_mn, b_mn=b_mn)
    want = (A.double() @ B.double().t()).float()
    assert torch.equal(got, want), f'max |diff| = {(got - want).abs().max().item()}'


def test_pair_epilogues():
    gen = torch.Generator(device='cuda').manual_seed(3)
    M, N, K = 20000, 384, 96
    A, B = _ints((M, K), gen)*0.25, _ints((N, K), gen)*0.25
    bias, gamma, res = (torch.randn(s, generator=gen, device='cuda') for s in ((N,), (N,), (M, N)))
    aux = torch.empty(M, N, device='cuda')
    got = F_.gemm_tf32(A, B, bias=bias, act='gelu', aux=aux, gamma=gamma, res=res)
    z = (A.double() @ B.double().t()) + bias.double()
    want = torch.nn.functional.gelu(z)*gamma.double() + res.double()
    assert (aux.double() - z).abs().max() < 1e-5
    assert (got.double() - want).abs().max() < 2e-5
    cs = torch.ones(N, device='cuda')
    Ai, Bi = _ints((M, K), gen), _ints((N, K), gen)
    got = F_.gemm_tf32(Ai, Bi, colsum=cs)
    want = Ai.double() @ Bi.double().t()
    assert torch.equal(got, want.float()) and torch.equal(cs, (want.sum(0) + 1).float())


def test_pair_split_k_accumulate():
    gen = torch.Generator(device='cuda').manual_seed(17)
    M, N, K = 1536, 384, 7680
    A, B, As, Bs = _operands(M, N, K, True, True, gen, ints=True)
    out = torch.ones(M, N, device='cuda')
    F_.gemm_tf32(As, Bs, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=8)
    want = (A.double() @ B.double().t()).float() + 1
    assert torch.equal(out, want)
