"""Developer aid: time the depthwise 7x7 / LayerNorm kernels at the four ConvNeXt-T stage shapes of BASELINE config 3.
  python tools/bench_dw.py [--b 8] [--only 0,2]"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import functional as F_

ap = argparse.ArgumentParser(); ap.add_argument('--b', type=int, default=8); ap.add_argument('--only', default=''); a = ap.parse_args()
dev = 'cuda'


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n


print(f'{"stage":6s} {"MB":>7s} | dw fwd   dw bwd(dgrad+wgrad)   ln fwd   ln bwd  (ms)  | ideal 2-pass ms')
for st, (C, h, w) in enumerate([(96, 96, 160), (192, 48, 80), (384, 24, 40), (768, 12, 20)]):
    if a.only and str(st) not in a.only.split(','): continue
    x = torch.randn(a.b, h, w, C, device=dev, requires_grad=True)
    wt = torch.randn(C, 1, 7, 7, device=dev, requires_grad=True)
    bs = torch.randn(C, device=dev, requires_grad=True)
    g, be = torch.randn(C, device=dev, requires_grad=True), torch.randn(C, device=dev, requires_grad=True)
    gy = torch.randn(a.b, h, w, C, device=dev)
    mb = x.numel()*4/1e6
    t_df = timeit(lambda: F_.dwconv7(x, wt, bs))
    t_dfb = timeit(lambda: F_.dwconv7(x, wt, bs).backward(gy))
    t_lf = timeit(lambda: F_.layer_norm(x, g, be))
    t_lfb = timeit(lambda: F_.layer_norm(x, g, be).backward(gy))
    print(f'st{st}    {mb:7.1f} | {t_df:7.3f} {t_dfb - t_df:9.3f}            {t_lf:7.3f} {t_lfb - t_lf:8.3f}        | {2*mb/6.5e3:7.3f}')
