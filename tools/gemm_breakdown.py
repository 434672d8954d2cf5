"""Developer aid: per-shape device time of every tensor-core call (stv_gemm_tf32 / stv_conv_*) in one eager training step.
  python tools/gemm_breakdown.py [--b 8] [--top 60]"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import functional as F_, synthetic as syn
from slowtv_monodepth_b200.optim import FlatAdamW
from slowtv_monodepth_b200.trainer import MonoDepthStep, default_cfg

ap = argparse.ArgumentParser()
ap.add_argument('--b', type=int, default=8); ap.add_argument('--H', type=int, default=384); ap.add_argument('--W', type=int, default=640)
ap.add_argument('--depth', default='convnext_tiny'); ap.add_argument('--pose', default='resnet18'); ap.add_argument('--top', type=int, default=60)
a = ap.parse_args()
torch.manual_seed(0)
model = MonoDepthStep(default_cfg(a.depth, a.pose)).to('cuda').train().to(memory_format=torch.channels_last)
opt = FlatAdamW(model.nets)
batch = syn.make_batch(a.b, 2, (a.H, a.W), seed=0, device='cuda')

def step():
    opt.zero_grad()
    loss, _, _ = model.step(batch)
    loss.backward()
    opt.step()

for _ in range(3): step()
F_.enable_kernel_timing(True, detail=True)
R = 3
for _ in range(R): step()
torch.cuda.synchronize()
kt = F_.kernel_timings()
tot = sum(sum(v) for v in kt.values())/R
print(f'{len(kt)} distinct (entry, shape) pairs, {tot:.2f} ms per step inside timed libstv calls (eager; includes launch gaps)')
for k, v in sorted(kt.items(), key=lambda kv: -sum(kv[1]))[:a.top]:
    print(f'{sum(v)/R:8.3f} ms  x{len(v)//R:3d}  {min(v):7.3f} min  {k}')
