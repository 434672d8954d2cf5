"""Developer aid: in-kernel timeline of CTA 0 of the row-segment convolution (trace build, see tools/gemm_trace.py).
  STV_LIB=.../libstv_trace.so python tools/conv3_trace.py N,H,W,Cin,Cout ..."""
import ctypes as C, statistics, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import _lib as L, functional as F_
L.lib()
fn = C.CDLL(str(L.LIB_PATH)).stv_debug_conv3_trace
fn.argtypes, fn.restype = [C.POINTER(C.c_ulonglong), C.c_int], C.c_int
for spec in sys.argv[1:] or ['16,96,160,64,64']:
    N, H, W, Cin, Cout = (int(v) for v in spec.split(','))
    x = torch.randn(N, H, W, Cin, device='cuda')
    w = torch.randn(Cout, Cin, 3, 3, device='cuda').contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, device='cuda')
    for _ in range(3): y = F_.conv2d_nhwc(x, w, b, pad=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = F_.conv2d_nhwc(x, w, b, pad=1); e1.record(); torch.cuda.synchronize()
    buf = (C.c_ulonglong*1040)(); assert fn(buf, 1040) == 0
    t = list(buf); T0, T1, Cm, E0, E1, T2 = t[:6]
    P = [v for v in t[16:528] if v >= T0]; Fw = [v for v in t[528:1040] if v >= T0]
    d = [Fw[i + 1] - Fw[i] for i in range(len(Fw) - 1)]
    lead = [Fw[i] - P[i] for i in range(min(len(P), len(Fw)))]
    print(f'{spec}: {e0.elapsed_time(e1)*1e3:.1f} us; set-up {T1 - T0} | first FULL after {Fw[0] - T1} | stage period median {statistics.median(d):.0f} min {min(d)} max {max(d)} '
          f'| epilogue of the first tile {E1 - E0} | exit at {T2 - T0}; stages traced {len(Fw)}; load latency median {statistics.median(lead):.0f}')
    print('   periods:', ' '.join(str(v) for v in d[:48]))
    print('   producer issue gaps:', ' '.join(str(P[i + 1] - P[i]) for i in range(min(40, len(P) - 1))))
