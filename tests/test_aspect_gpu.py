"""GPU: the CUDA aspect-ratio augmentation (slowtv_monodepth_b200/aspect_ratio.py -> stv_resample_bilinear through the C ABI)
against the oracle port (oracle/aspect.py, float64) under identical seeds, against the reference-made golden file, and the
kernel's two modes against torch's own F.interpolate / F.grid_sample in float64. Tolerance: 5e-6 absolute on [0,1] images
(float32 sample positions and lerps, as ATen's own CUDA kernels compute them; measured worst case 2.4e-6)."""
import json
import random
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import aspect as OA
from oracle.make_golden_aspect import CASES, make_batch
from slowtv_monodepth_b200 import aspect_ratio as AR, functional as F_

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent/'golden'/'aspect_cases.json').read_text())
TOL = 5e-6
# White-noise images (neighbouring pixels differ by O(1)): a float32 sample position near column 600 carries ulp ~6e-5, which the
# lerp turns into the same absolute error — inherent to float32 positions (ATen's / the reference's own path included).
TOL_NOISE = 1e-4


def _to(batch, dev, dt):
    x, y, m = batch
    mv = lambda d: {k: (v.to(dev, dt) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
    return mv(x), mv(y), dict(m)


@pytest.mark.parametrize('i', range(len(CASES)))
def test_aug_matches_oracle_and_golden(i):
    seed, b, n, shape, p, cmin, cmax, ref = CASES[i]
    random.seed(seed); torch.manual_seed(seed)
    want = OA.aspect_ratio_aug(make_batch(seed, b, n, shape), p=p, crop_min=cmin, crop_max=cmax, ref_shape=ref)
    random.seed(seed); torch.manual_seed(seed)
    got = AR.aspect_ratio_aug(_to(make_batch(seed, b, n, shape), 'cuda', torch.float32), p=p, crop_min=cmin, crop_max=cmax, ref_shape=ref)
    assert got[2].get('augs', []) == GOLD['results'][i]['augs']
    for part in (0, 1):
        for k in ('imgs', 'supp_imgs'):
            a, w = got[part][k].cpu().double(), want[part][k]
            assert a.shape == w.shape, (k, a.shape, w.shape)
            scale = 1.0 if part == 1 else 1/0.224   # x.* are standardised: errors scale with 1/std
            assert (a - w).abs().max().item() < TOL*scale*2, (part, k, (a - w).abs().max().item())
    assert torch.allclose(got[1]['K'].cpu().double(), want[1]['K'], rtol=1e-6)
    g = GOLD['results'][i]
    assert list(got[0]['imgs'].shape[-2:]) == g['shape']
    assert abs(got[1]['imgs'].double().sum().item() - g['sum']['y.imgs']) < 1e-5*abs(g['sum']['y.imgs']) + 1e-3


@pytest.mark.parametrize('shape,size', [((2, 3, 37, 61), (64, 96)), ((1, 3, 96, 160), (32, 64)), ((3, 1, 40, 72), (40, 72)), ((1, 2, 5, 7), (9, 3))])
def test_interp_mode_matches_torch(shape, size):
    x = torch.rand(shape, device='cuda')
    got = F_.resample_bilinear(x, size, mode='interp')
    want = F.interpolate(x.double(), size=size, mode='bilinear', align_corners=False)
    assert (got.double() - want).abs().max().item() < TOL_NOISE


@pytest.mark.parametrize('src,dst', [((96, 160), (53, 127)), ((97, 161), (59, 35)), ((384, 640), (259, 518))])
def test_grid_mode_matches_grid_sample(src, dst):
    x = torch.rand(2, 3, *src, device='cuda')
    got = AR.center_crop(x, dst)
    want = OA.center_crop(x.double().cpu(), dst)
    assert got.shape == want.shape
    assert (got.double().cpu() - want).abs().max().item() < TOL_NOISE


def test_five_dimensional_support_frames_and_errors():
    x = torch.rand(2, 2, 3, 33, 47, device='cuda')
    got = F_.resample_bilinear(x, (32, 64), mode='interp')
    want = F.interpolate(x.flatten(0, 1), size=(32, 64), mode='bilinear', align_corners=False).unflatten(0, (2, 2))
    assert (got - want).abs().max().item() < TOL_NOISE
    with pytest.raises(ValueError): F_.resample_bilinear(x, (0, 4), mode='interp')
    with pytest.raises(ValueError): F_.resample_bilinear(x, (4, 4), mode='nearest')
    with pytest.raises(Exception): F_.resample_bilinear(x.cpu(), (4, 4), mode='interp')
