"""Developer aid: kernel-level time breakdown of one training step with torch.profiler (cheap; no ncu replay).
  python tools/step_profile.py [--b 8] [--depth convnext_tiny] [--pose resnet18]"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile
from slowtv_monodepth_b200 import synthetic as syn
from slowtv_monodepth_b200.optim import FlatAdamW
from slowtv_monodepth_b200.trainer import MonoDepthStep, default_cfg

ap = argparse.ArgumentParser()
ap.add_argument('--b', type=int, default=8); ap.add_argument('--H', type=int, default=384); ap.add_argument('--W', type=int, default=640)
ap.add_argument('--depth', default='convnext_tiny'); ap.add_argument('--pose', default='resnet18'); ap.add_argument('--top', type=int, default=45)
a = ap.parse_args()
torch.backends.cudnn.benchmark = True
torch.set_float32_matmul_precision('high')
dev = 'cuda'
torch.manual_seed(0)
model = MonoDepthStep(default_cfg(a.depth, a.pose)).to(dev).train().to(memory_format=torch.channels_last)
opt = FlatAdamW(model.nets)
batch = syn.make_batch(a.b, 2, (a.H, a.W), seed=0, device=dev)

def step():
    opt.zero_grad()
    loss, _, _ = model.step(batch)
    loss.backward()
    opt.step()

for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type is not None]
tot = {}
for e in prof.events():
    if e.device_type is not None and str(e.device_type).endswith('CUDA'):
        tot[e.name] = tot.get(e.name, [0, 0]); tot[e.name][0] += e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total; tot[e.name][1] += 1
T = sum(v[0] for v in tot.values())
print(f'{len(tot)} distinct kernels, {T/2e3:.2f} ms of kernel time per step')
for name, (t, c) in sorted(tot.items(), key=lambda x: -x[1][0])[:a.top]:
    print(f'{t/2e3:8.3f} ms {100*t/T:5.1f}%  x{c//2:4d}  {name[:120]}')
