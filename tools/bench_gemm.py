"""Developer aid: time the six tcgen05 GEMMs of one ConvNeXt block (fwd fc1/fc2, bwd dz/dx/dW1/G) per stage of ConvNeXt-T at
BASELINE config 3, against torch (cuBLAS TF32).  python tools/bench_gemm.py [--b 8]"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import functional as F_

ap = argparse.ArgumentParser(); ap.add_argument('--b', type=int, default=8); ap.add_argument('--only', default=''); ap.add_argument('--once', action='store_true', help='run every case once, ours only (ncu captures)'); a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = True
dev = 'cuda'
torch.cuda.set_stream(torch.cuda.Stream())


def timeit(fn, n=10):
    """GPU time per call: n calls captured in ONE CUDA graph and replayed, so that the host's launch path (ctypes + tensor-map
    encoding, ~20-30 us per call — longer than most of these kernels) is not in the measurement. Operands are re-used by the n
    calls, i.e. L2-warm where they fit: the in-step numbers (tools/step_profile.py) are the cold-cache counterpart."""
    st = torch.cuda.current_stream()   # the script runs on ONE non-default stream (set below): tensors, warm-up and capture alike —
    for _ in range(3): fn()            # autograd's AccumulateGrad nodes stay bound to the stream their tensor was created on
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    best = float('inf')
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)/n)
    return best


print(f'{"gemm":22s} {"M":>7s} {"N":>5s} {"K":>7s} | ours ms   TF/s   GB/s | cublas ms')
for st, (C, hw) in enumerate([(96, 96*160), (192, 48*80), (384, 24*40), (768, 12*20)]):
    if a.only and str(st) not in a.only.split(','): continue
    M, Hd = a.b*hw, 4*C
    x, res, g = (torch.randn(M, C, device=dev) for _ in range(3))
    z, h, dz = (torch.randn(M, Hd, device=dev) for _ in range(3))
    w1, w2 = torch.randn(Hd, C, device=dev), torch.randn(C, Hd, device=dev)
    b1, b2, gam = torch.randn(Hd, device=dev), torch.randn(C, device=dev), torch.randn(C, device=dev)
    dw1, G = torch.zeros_like(w1), torch.zeros_like(w2)
    hout, out = torch.empty_like(h), torch.empty_like(x)
    cases = [
        ('fc1+gelu', M, Hd, C, lambda: F_.gemm_tf32(x, w1, bias=b1, act='gelu', aux=z, out=hout), lambda: torch.nn.functional.gelu(torch.addmm(b1, x, w1.t())), 4*(M*C + 2*M*Hd)),
        ('fc2+scale+res', M, C, Hd, lambda: F_.gemm_tf32(h, w2, bias=b2, gamma=gam, res=res, out=out), lambda: res + gam*torch.addmm(b2, h, w2.t()), 4*(M*Hd + 2*M*C)),
        ('dz=g.W2*gelu\'', M, Hd, C, lambda: F_.gemm_tf32(g, w2, b_mn=True, dact='gelu', dact_src=z, out=hout), lambda: torch.ops.aten.gelu_backward(g @ w2, z), 4*(M*C + 2*M*Hd)),
        ('dx=dz.W1', M, C, Hd, lambda: F_.gemm_tf32(dz, w1, b_mn=True, out=out), lambda: dz @ w1, 4*(M*Hd + M*C)),
        ('dW1=dz^T.x', Hd, C, M, lambda: F_.gemm_tf32(dz, x, a_mn=True, b_mn=True, out=dw1, accumulate=True, split_k=F_._split_k(Hd, C, M)), lambda: dz.t() @ x, 4*(M*Hd + M*C)),
        ('G=g^T.h', C, Hd, M, lambda: F_.gemm_tf32(g, h, a_mn=True, b_mn=True, out=G, accumulate=True, split_k=F_._split_k(C, Hd, M)), lambda: g.t() @ h, 4*(M*Hd + M*C)),
    ]
    for name, m, n, k, ours, ref, nbytes in cases:
        if a.once:
            ours(); torch.cuda.synchronize(); print(f'st{st} {name}'); continue
        t, tr = timeit(ours), timeit(ref)
        print(f'st{st} {name:18s} {m:7d} {n:5d} {k:7d} | {t:7.3f} {2*m*n*k/t/1e9:6.1f} {nbytes/t/1e6:6.0f} | {tr:7.3f}')
