"""Developer aid (not a test): where does the per-pixel gradient error of the single-pass kernel sit at a full-size case?
  python tools/debug_fullsize.py [b n S H W seed]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
from slowtv_monodepth_b200 import synthetic as syn
from tests import util as U

a = [int(v) for v in sys.argv[1:]] + [2, 2, 4, 384, 640, 11][len(sys.argv) - 1:]
b, n, S, H, W, seed = a
inp = syn.make_loss_inputs(b, n, S, (H, W), seed=seed)
cfg = dict(b=b, n=n, S=S, shape=(H, W))
got = U.run_cuda(inp, cfg)
torch.cuda.synchronize()
sel = got['sel'].cpu()
with U.eps32():
    bad, _ = U.unstable_pixels(inp, cfg)
    want = U.run_oracle(inp, cfg, torch.float64, forced_sel=sel)
    ref32 = U.run_oracle(inp, cfg, torch.float32, forced_sel=sel)
for s in range(S):
    good = ~(F.max_pool2d(bad[s*b:(s + 1)*b].float(), 3, 1, 1) > 0)
    g, w, r = (x.detach().double().cpu()*good for x in (got[f'g_up{s}'], want[f'g_dispup{s}'], ref32[f'g_dispup{s}']))
    e2, f2 = (g - w).pow(2).flatten(), (r - w).pow(2).flatten()
    tot, nrm = e2.sum(), w.pow(2).sum()
    print(f'scale {s}: ours {(tot/nrm).sqrt():.3e}  float32 oracle {(f2.sum()/nrm).sqrt():.3e}  compared {good.float().mean():.1%}')
    srt = e2.sort(descending=True)
    for k in (1, 10, 100, 1000, 10000):
        rest = (tot - srt[0][:k].sum()).clamp(min=0)
        print(f'   without the worst {k:6d} pixels: {(rest/nrm).sqrt():.3e}   (float32 oracle without ITS worst {k}: {((f2.sum() - f2.sort(descending=True)[0][:k].sum()).clamp(min=0)/nrm).sqrt():.3e})')
    for i in srt[1][:12]:
        bi, _, y, x = [int(v) for v in torch.unravel_index(i, g.shape)]
        print(f'      img {bi} y {y:4d} x {x:4d} (x%28={x % 28:2d}) got {g.flatten()[i]:+.5e} want {w.flatten()[i]:+.5e} f32 {r.flatten()[i]:+.5e} sel {int(sel[s*b + bi, 0, y, x])}')
