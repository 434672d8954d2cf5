"""CPU: the oracle port of the aspect-ratio augmentation (oracle/aspect.py) against golden answers produced by the reference's
own `aspect_ratio_aug` (oracle/make_golden_aspect.py; kornia's crop kernel is the restated one — see oracle/aspect.py), plus the
host-side sampling logic of the product module (pure Python/torch-CPU) against the same goldens."""
import json
import random
from pathlib import Path

import pytest
import torch

from oracle import aspect as OA
from oracle.make_golden_aspect import CASES, run
from slowtv_monodepth_b200 import aspect_ratio as AR

GOLD = json.loads((Path(__file__).parent/'golden'/'aspect_cases.json').read_text())


@pytest.mark.parametrize('i', range(len(CASES)))
def test_oracle_matches_reference_golden(i):
    got, want = run(OA.aspect_ratio_aug, CASES[i]), GOLD['results'][i]
    assert got['augs'] == want['augs'] and got['shape'] == want['shape']
    assert torch.allclose(torch.tensor(got['K']), torch.tensor(want['K']), rtol=1e-12, atol=0)
    for k, v in want['sum'].items(): assert abs(got['sum'][k] - v) <= 1e-9*max(1., abs(v)), k
    for k in ('probe_x_imgs', 'probe_y_supp'):
        assert torch.allclose(torch.tensor(got[k], dtype=torch.float64), torch.tensor(want[k], dtype=torch.float64), rtol=1e-10, atol=1e-12), k


@pytest.mark.parametrize('i', range(len(CASES)))
def test_product_sampling_logic_matches_reference_golden(i):
    """Same seeds -> same crop / resize decisions as the reference (the strings it logs into m['augs'])."""
    seed, b, n, shape, p, cmin, cmax, ref = CASES[i]
    want = GOLD['results'][i]['augs']
    random.seed(seed); torch.manual_seed(seed)
    augs, sh = [], tuple(shape)
    if random.random() > p:
        if ref and tuple(ref) != sh: augs.append(str(AR.sample_resize(sh, ref, eps=1)))
    else:
        crop, ratio = AR.sample_crop(sh, cmin, cmax)
        augs.append(f'{list(sh)} -> {crop} -> {AR.LABELS[AR.RATIOS.index(ratio)]}')
        augs.append(str(AR.sample_resize(crop, ref or sh, eps=0.8)))
    assert augs == want


def test_resize_sizes_are_multiples_of_32_within_budget():
    for shape in [(53, 127), (300, 168), (384, 640), (97, 161)]:
        out = AR.sample_resize(shape, (384, 640), eps=0.8)
        assert all(s % 32 == 0 and s > 0 for s in out) and out[0]*out[1] <= 0.8*384*640


def test_crop_positions_match_the_grid_sample_formulation():
    """`_crop_axis` (what the CUDA kernel evaluates) vs. the positions F.affine_grid/grid_sample see in the oracle's restatement."""
    for src, dst in [(96, 53), (160, 127), (97, 59), (161, 35)]:
        a, b = AR._crop_axis(src, dst)
        start = OA.crop_start(src, dst)
        j = torch.arange(dst, dtype=torch.float64)
        want = (start + (j + 0.5)*(dst - 1)/dst)*src/(src - 1) - 0.5
        assert torch.allclose(a*j + b, want, rtol=0, atol=1e-9)


def test_host_tensors_raise():
    x = torch.zeros(1, 3, 8, 8)
    with pytest.raises(Exception): AR.center_crop(x, (4, 4))
