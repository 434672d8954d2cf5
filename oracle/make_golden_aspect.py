"""Generate tests/golden/aspect_cases.json by executing the REAL reference `aspect_ratio_aug` (src/core/aspect_ratio.py, from
/root/reference) on CPU under fixed seeds.  TEST INFRASTRUCTURE; run in the build container only:  python oracle/make_golden_aspect.py

kornia (0.6.10 pinned by the reference) is not installed: its `center_crop` is supplied by the restatement in oracle/aspect.py, so
these fixtures pin the reference's sampling logic, resize, intrinsics handling and batch plumbing — everything except kornia's own
kernel (oracle/aspect.py header: parity unpinned for that function).
"""
from __future__ import annotations

import json
import random
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import aspect as OA, ref_shim  # noqa: E402

CASES = [  # (seed, b, n, (H, W), p, crop_min, crop_max, ref_shape)
    (0, 2, 2, (96, 160), 1.0, 0.5, 1.0, None),
    (1, 1, 2, (128, 192), 1.0, 0.5, 1.0, (96, 160)),
    (2, 2, 1, (96, 160), 0.0, 0.5, 1.0, (64, 128)),     # p-miss: plain resize to ref_shape
    (3, 1, 4, (120, 200), 0.7, 0.6, 0.9, None),
    (4, 1, 2, (96, 160), 0.0, 0.5, 1.0, None),          # p-miss without ref_shape: untouched
    (5, 2, 2, (192, 320), 1.0, 0.5, 1.0, (192, 320)),
    (6, 1, 2, (97, 161), 1.0, 0.5, 1.0, (96, 160)),     # odd source size
    (7, 1, 2, (96, 160), 1.0, 0.5, 1.0, None),
]


def make_batch(seed: int, b: int, n: int, shape):
    g = torch.Generator().manual_seed(1000 + seed)
    H, W = shape
    base = lambda *s: torch.nn.functional.interpolate(torch.rand(*s, H//8 + 1, W//8 + 1, generator=g, dtype=torch.float64), size=(H, W),
                                                      mode='bilinear', align_corners=True)
    y = {'imgs': base(b, 3), 'supp_imgs': base(n*b, 3).unflatten(0, (n, b)),
         'K': torch.tensor([[.58*W, 0, .5*W, 0], [0, 1.92*H, .5*H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64).expand(b, 4, 4).clone()}
    mean, std = torch.tensor([.485, .456, .406], dtype=torch.float64).view(3, 1, 1), torch.tensor([.229, .224, .225], dtype=torch.float64).view(3, 1, 1)
    x = {'imgs': (y['imgs'] - mean)/std, 'supp_imgs': (y['supp_imgs'] - mean)/std, 'supp_idxs': torch.tensor([-1, 1, -2, 2][:n])}
    return x, y, {}


def summary(batch) -> dict:
    x, y, m = batch
    probe = lambda t: t.flatten()[:: max(1, t.numel()//64)][:64].tolist()
    return {'augs': m.get('augs', []), 'shape': list(x['imgs'].shape[-2:]), 'K': y['K'].tolist(),
            'sum': {f'{k}.{k2}': float(d[k2].sum()) for k, d in (('x', x), ('y', y)) for k2 in ('imgs', 'supp_imgs')},
            'probe_x_imgs': probe(x['imgs']), 'probe_y_supp': probe(y['supp_imgs'])}


def run(fn, case) -> dict:
    seed, b, n, shape, p, cmin, cmax, ref = case
    batch = make_batch(seed, b, n, shape)
    random.seed(seed); torch.manual_seed(seed)
    return summary(fn(batch, p=p, crop_min=cmin, crop_max=cmax, ref_shape=ref))


def reference_fn():
    ref_shim.load()
    import kornia.geometry.transform as KT  # the placeholder module installed by ref_shim
    KT.center_crop = lambda t, size, mode='bilinear', align_corners=False: OA.center_crop(t, size)
    import importlib
    ar = importlib.import_module('src.core.aspect_ratio')
    ar.KT.center_crop = KT.center_crop
    return ar.aspect_ratio_aug


if __name__ == '__main__':
    fn = reference_fn()
    out = {'cases': [list(c[:3]) + [list(c[3])] + list(c[4:7]) + [list(c[7]) if c[7] else None] for c in CASES],
           'results': [run(fn, c) for c in CASES]}
    path = ROOT/'tests'/'golden'/'aspect_cases.json'
    path.write_text(json.dumps(out))
    print(f'wrote {path} ({path.stat().st_size} bytes)')
    for c, r in zip(CASES, out['results']): print(c[0], r['augs'], r['shape'])
