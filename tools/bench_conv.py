"""Developer aid: time the tcgen05 implicit-GEMM convolutions against cuDNN (torch, TF32, channels-last) at the layer shapes
of BASELINE config 3.  python tools/bench_conv.py [--b 8]"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
from slowtv_monodepth_b200 import functional as F_

ap = argparse.ArgumentParser(); ap.add_argument('--b', type=int, default=8); ap.add_argument('--only', default=''); a = ap.parse_args()
torch.backends.cudnn.benchmark = True; torch.backends.cudnn.allow_tf32 = True
dev = 'cuda'
torch.cuda.set_stream(torch.cuda.Stream())
#         name            H    W    C1   C2  up1   Cout R st pad reflect
SHAPES = [('upconv_4_0', 12, 20, 768, 0, False, 256, 3, 1, 1, True),
          ('upconv_4_1', 24, 40, 256, 384, True, 256, 3, 1, 1, True),
          ('upconv_3_1', 48, 80, 128, 192, True, 128, 3, 1, 1, True),
          ('upconv_2_1', 96, 160, 64, 96, True, 64, 3, 1, 1, True),
          ('upconv_1_0', 96, 160, 64, 0, False, 32, 3, 1, 1, True),
          ('upconv_1_1', 192, 320, 32, 0, True, 32, 3, 1, 1, True),
          ('upconv_0_0', 192, 320, 32, 0, False, 16, 3, 1, 1, True),
          ('upconv_0_1', 384, 640, 16, 0, True, 16, 3, 1, 1, True),
          ('outconv_0', 384, 640, 16, 0, False, 1, 3, 1, 1, True),
          ('res_l1', 96, 160, 64, 0, False, 64, 3, 1, 1, False),
          ('res_l2_s2', 96, 160, 64, 0, False, 128, 3, 2, 1, False),
          ('res_l3', 24, 40, 256, 0, False, 256, 3, 1, 1, False),
          ('res_l4', 12, 20, 512, 0, False, 512, 3, 1, 1, False),
          ('cnx_down1', 96, 160, 96, 0, False, 192, 2, 2, 0, False),
          ('res_conv1', 384, 640, 8, 0, False, 64, 7, 2, 3, False)]


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


print(f'{"layer":12s} {"GF(fwd)":>8s} | ours fwd  bwd (ms) TF/s(fwd) | cudnn fwd  bwd (ms)')
for name, H, W, C1, C2, up1, Cout, R, st, pad, refl in SHAPES:
    if a.only and name not in a.only.split(','): continue
    N = a.b*(2 if name.startswith('res') else 1)
    s1 = torch.randn((N, H//2, W//2, C1) if up1 else (N, H, W, C1), device=dev).requires_grad_()
    s2 = torch.randn(N, H, W, C2, device=dev).requires_grad_() if C2 else None
    w = torch.randn(Cout, C1 + C2, R, R, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    b = torch.randn(Cout, device=dev).requires_grad_()
    kw = dict(up1=up1, stride=st, pad=pad, reflect=refl)
    y = F_.conv2d_nhwc(s1, w, b, src2=s2, act='elu', **kw)
    dA = torch.randn_like(y)
    P, Q = y.shape[1:3]
    gf = 2*N*P*Q*Cout*R*R*(C1 + C2)/1e9
    t_f = timeit(lambda: F_.conv2d_nhwc(s1, w, b, src2=s2, act='elu', **kw))
    def ours_fb():
        F_.conv2d_nhwc(s1, w, b, src2=s2, act='elu', **kw).backward(dA)
    t_fb = timeit(ours_fb)

    x1 = s1.detach().permute(0, 3, 1, 2).requires_grad_()
    x2 = s2.detach().permute(0, 3, 1, 2).requires_grad_() if C2 else None
    def ref():
        x = F.interpolate(x1, scale_factor=2, mode='nearest') if up1 else x1
        if x2 is not None: x = torch.cat([x, x2], 1)
        p = pad
        if refl: x, p = F.pad(x, (pad,)*4, mode='reflect'), 0
        return F.elu(F.conv2d(x, w, b, st, p))
    r_f = timeit(ref)
    dAr = dA.permute(0, 3, 1, 2)
    r_fb = timeit(lambda: ref().backward(dAr))
    print(f'{name:12s} {gf:8.2f} | {t_f:8.3f} {t_fb - t_f:8.3f} {gf/t_f:8.1f}    | {r_f:8.3f} {r_fb - r_f:8.3f}')
