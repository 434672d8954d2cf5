"""Developer aid (not a test): print per-output parity of the CUDA loss path vs the float64 oracle, with the location of
the worst pixels. Usage on the GPU box:  python tools/debug_parity.py [case ...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests import util as U
from oracle import loss as OL


def worst(name, a, b, k=5):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = (a - b).abs()
    print(f'  {name:12s} rel={U.rel(a, b):.3e} max_abs={d.max().item():.3e} ref_max={b.abs().max().item():.3e}')
    if d.ndim == 4 and d.max() > 0:
        idx = d.flatten().topk(k)[1]
        for i in idx:
            pos = torch.unravel_index(i, d.shape)
            print('      at', [int(p) for p in pos], f'got={a.flatten()[i].item():.6e} want={b.flatten()[i].item():.6e}')


def vs_case(name):
    from slowtv_monodepth_b200.geometry import ViewSynth
    inp, cfg, _ = U.load_golden(name)
    d64 = U.cast(inp, torch.float64)
    H, W = cfg['shape']
    feat = torch.rand(cfg['b'], 5, H, W, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    mn, mx = cfg.get('min_depth', 0.1), cfg.get('max_depth', 100.)
    depth = OL.disp_to_depth(OL.resize_bilinear(d64['disps'][0], (H, W)), mn, mx)
    T = OL.T_from_AAt(d64['aa'][0], d64['t'][0])

    def run(fn, dt, dev):
        x, dp, Tm, K = (v.to(dev, dt).clone().requires_grad_() for v in (feat, depth, T, d64['K']))
        w, dw, valid = fn(x, dp, Tm, K)
        ((w*w).sum()).backward()
        return w.detach(), dw.detach(), valid, x.grad, dp.grad, Tm.grad, K.grad
    want = run(OL.view_synth, torch.float64, 'cpu')
    got = run(ViewSynth((H, W)), torch.float32, 'cuda')
    print(f'== view_synth on {name}')
    for j, nm in ((0, 'warp'), (1, 'dwarp'), (3, 'g_input'), (4, 'g_depth'), (5, 'g_T'), (6, 'g_K')):
        worst(nm, got[j], want[j])


def loss_case(name):
    inp, cfg, ref = U.load_golden(name)
    got = U.run_cuda(inp, cfg)
    sel = got['sel'].cpu()
    fsel = sel if cfg.get('use_min', True) else None
    with U.eps32():
        want = U.run_oracle(inp, cfg, torch.float64, forced_sel=fsel)
        bad, _ = U.unstable_pixels(inp, cfg)
    print(f'== loss stack on {name}  (unstable pixels: {bad.float().mean().item():.3%})')
    b = cfg['b']
    for k in sorted(got):
        if k in ('sel', 'automask0'): continue
        worst(k, got[k], want[k])
        if k.startswith('g_depth'):
            s = int(k[-1]); good = ~(torch.nn.functional.max_pool2d(bad[s*b:(s + 1)*b].float(), 3, 1, 1) > 0)
            print(f'      masked rel = {U.rel_masked(got[k], want[k], good):.3e}')
            worst(k + '*', got[k].cpu()*good, want[k]*good)


if __name__ == '__main__':
    names = sys.argv[1:] or U.LOSS_CASES
    for n in names:
        vs_case(n) if n.startswith('vs:') is False and False else None
    for n in names:
        if n.startswith('vs:'): vs_case(n[3:])
        else: loss_case(n)
