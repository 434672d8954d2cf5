"""Developer aid: run the fused loss stack alone at a benchmark shape (for ncu captures and quick timing).
  python tools/profile_loss.py [--b 8 --H 384 --W 640 --n 2 --S 4 --iters 5]"""
import argparse
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import geometry as G, handlers as Hd, synthetic as syn, functional as F_
from slowtv_monodepth_b200.losses import ReconstructionLoss
from slowtv_monodepth_b200.regularizers import SmoothReg

ap = argparse.ArgumentParser()
for k, v in dict(b=8, H=384, W=640, n=2, S=4, iters=5).items(): ap.add_argument(f'--{k}', type=int, default=v)
ap.add_argument('--mode', default='disp', choices=['disp', 'depth', 'two-pass'])
ap.add_argument('--kernels', action='store_true', help='also print the CUPTI duration of every kernel of one iteration (torch.profiler)')
a = ap.parse_args()
dev = 'cuda'
d = syn.make_loss_inputs(a.b, a.n, a.S, (a.H, a.W), seed=0)
d = {k: ([x.to(dev) for x in v] if isinstance(v, list) else v.to(dev)) for k, v in d.items()}
disps = [x.requires_grad_() for x in d['disps']]
aa, t = d['aa'].requires_grad_(), d['t'].requires_grad_()
crit, sm = ReconstructionLoss('ssim', True, True), SmoothReg(use_edges=True)
F_.enable_kernel_timing(True)
prof = None
for it in range(a.iters + (1 if a.kernels else 0)):
    if a.kernels and it == a.iters:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        prof = profile(activities=[ProfilerActivity.CUDA]); prof.__enter__()
    Ts = G.T_from_AAt(aa, t)
    F_.PHOTO_FORCE_TWO_PASS = a.mode == 'two-pass'
    if a.mode == 'disp':
        l1, _ = Hd.image_recon(crit, None, None, None, d['imgs'], d['supp_imgs'], Ts, d['K'], want_warp=False, disps=dict(enumerate(disps)),
                               depth_range=(0.1, 100.))
    else:
        depths = {s: G.upsample_to_depth(x, (a.H, a.W), 0.1, 100.)[1] for s, x in enumerate(disps)}
        for x in depths.values(): x._stv_src = None if a.mode == 'depth' else x._stv_src
        l1 = crit.fused(list(depths.values()), d['imgs'], d['supp_imgs'], Ts, d['K'], want_warp=False)[0]
    l2, _ = Hd.disp_smooth(sm, dict(enumerate(disps)), d['imgs'], want_maps=False)
    (l1 + 1e-3*l2).backward()
torch.cuda.synchronize()
if prof is not None:
    prof.__exit__(None, None, None)
    for e in sorted(prof.events(), key=lambda e: e.time_range.start):
        if str(e.device_type).endswith('CUDA'): print(f'   {e.device_time_total:9.1f} us  {e.name[:110]}')
px = a.b*a.H*a.W
tot = 0.
for k, v in F_.kernel_timings().items():
    v = v[1:] if len(v) > 1 else v
    ms = sum(v)/len(v)
    tot += ms if k.startswith('stv_photo') else 0.
    print(f'{k:16s} {ms*1e3:9.1f} us')
by = (12 + 12*a.n + 12*a.n*a.S + 4*a.S)*px + (12 + 12*a.n + 12*a.n*a.S + 8*a.S)*px   # forward + backward algorithmic bytes
print(f'photometric pair ({a.mode}): {tot*1e3:.1f} us for {by/1e6:.0f} MB algorithmic = {by/1e9/(tot/1e3):.0f} GB/s')
print('loss', l1.item(), l2.item())
