"""CPU: host-side bookkeeping of the step mirror (no kernels): which transform each support frame gets (src/core/trainer.py:347)."""
import pytest
import torch

from slowtv_monodepth_b200.trainer import MonoDepthStep


def test_stereo_support_takes_the_calibrated_baseline():
    x = {'imgs': torch.zeros(1, 3, 8, 8)}
    fwd = {'_idxs': [-1, 0, 1], 'T_-1': torch.eye(4)[None], 'T_1': 3*torch.eye(4)[None], 'disp': {}}
    out = MonoDepthStep.forward_postprocess(None, fwd, x, {'T_stereo': 2*torch.eye(4)[None]}, want_up=False)
    assert out['Ts'].shape == (3, 1, 4, 4)
    assert [out['Ts'][k, 0, 0, 0].item() for k in range(3)] == [1., 2., 3.]


@pytest.mark.parametrize('idxs,y', [([0], {}), ([2], {'T_stereo': torch.eye(4)[None]})])
def test_missing_transforms_raise_a_clear_error(idxs, y):
    with pytest.raises(KeyError): MonoDepthStep.forward_postprocess(None, {'_idxs': idxs, 'disp': {}}, {'imgs': torch.zeros(1, 3, 8, 8)}, y, want_up=False)
