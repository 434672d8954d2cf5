"""Developer aid: device time and achieved HBM GB/s of the aspect-ratio augmentation at BASELINE config 3 (b=8, n=2, 384x640).
  python tools/bench_aspect.py"""
import random, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from slowtv_monodepth_b200 import aspect_ratio as AR, synthetic as syn

b, n, shape = 8, 2, (384, 640)
base = syn.make_batch(b, n, shape, seed=0, device='cuda')
def fresh():
    x, y, m = base
    return ({k: v.clone() if torch.is_tensor(v) else v for k, v in x.items()}, {k: v.clone() if torch.is_tensor(v) else v for k, v in y.items()}, {})
for seed in range(3):
    random.seed(seed); torch.manual_seed(seed)
    bt = fresh()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = AR.aspect_ratio_aug(bt, p=1.0, ref_shape=shape); e1.record(); torch.cuda.synchronize()
    crop = eval(out[2]['augs'][0].split(' -> ')[1]); res = eval(out[2]['augs'][1])
    planes = 2*(1 + n)*b*3
    # algorithmic bytes: read the crop window once, write + re-read the crop, write the result
    nbytes = planes*4*(crop[0]*crop[1]*3 + res[0]*res[1])
    ms = e0.elapsed_time(e1)
    print(f'seed {seed}: {out[2]["augs"]}  {ms*1e3:7.1f} us (incl. host launch gaps)  {nbytes/ms/1e6:7.1f} GB/s algorithmic')
