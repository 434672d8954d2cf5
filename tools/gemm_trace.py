"""Developer aid: in-kernel clock64 timeline of CTA 0 of the tcgen05 GEMM (needs a library built with -DSTV_GEMM_TRACE:
tools/build_trace_lib.sh -> slowtv_monodepth_b200/csrc/build/libstv_trace.so, loaded through STV_LIB).
  STV_LIB=slowtv_monodepth_b200/csrc/build/libstv_trace.so python tools/gemm_trace.py M N K [a_mn b_mn split_k] ..."""
import ctypes as C, statistics, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import _lib as L, functional as F_

lib = L.lib()
fn = C.CDLL(str(L.LIB_PATH)).stv_debug_gemm_trace
fn.argtypes, fn.restype = [C.POINTER(C.c_ulonglong), C.c_int], C.c_int


def run(M, N, K, a_mn=0, b_mn=0, split_k=1):
    A = torch.randn((K, M) if a_mn else (M, K), device='cuda')
    B = torch.randn((K, N) if b_mn else (N, K), device='cuda')
    out = torch.zeros(M, N, device='cuda')
    for _ in range(3): F_.gemm_tf32(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, accumulate=split_k > 1, split_k=split_k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); F_.gemm_tf32(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, accumulate=split_k > 1, split_k=split_k); e1.record()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong*1040)()
    assert fn(buf, 1040) == 0
    t = list(buf)
    T0, T1, Cm, E0, E1, T2 = t[:6]
    nkb = (K + 31)//32
    per_split = -(-nkb//split_k)
    P = [x for x in t[16:16 + 512] if x >= T0][:512]
    Fw = [x for x in t[528:528 + 512] if x >= T0][:512]
    n = min(len(Fw), per_split)
    d = [Fw[i + 1] - Fw[i] for i in range(n - 1)]
    lead = [Fw[i] - P[i] for i in range(min(n, len(P)))]
    print(f'M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} sk={split_k}: {e0.elapsed_time(e1)*1e3:.1f} us by events; CTA 0 (first tile, {n} k-blocks):')
    print(f'   set-up {T1 - T0} clk | first FULL after {Fw[0] - T1} | k-block period median {statistics.median(d) if d else 0:.0f} (min {min(d) if d else 0}, max {max(d) if d else 0}) '
          f'| last FULL -> accumulator seen by epilogue {E0 - Fw[n - 1]} | epilogue {E1 - E0} | exit at {T2 - T0} after entry')
    print(f'   load latency (producer issue -> MMA sees FULL): first {lead[0]}, median {statistics.median(lead):.0f}')
    print('   periods:', ' '.join(str(x) for x in d[:40]))


args = sys.argv[1:]
shapes = [a.split(',') for a in args] or [['7680', '384', '1536']]
for sh in shapes: run(*[int(v) for v in sh])
