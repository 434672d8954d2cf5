"""GPU: the fused flat-buffer AdamW kernel (stv_adamw_step, through FlatAdamW) against torch.optim.AdamW with timm's parameter
groups (biases and 1-D parameters get no weight decay — `param_groups_weight_decay`, what the reference's
`create_optimizer_v2(nets, 'adamw', ...)` builds, src/tools/parsers.py:205-243), five steps on the same gradients."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


class _Toy(nn.Module):
    """Every parameter kind of the real networks: 4-D filters (stored channels-last in the flat buffer), matrices, biases,
    1-D scales, odd element counts (the flat buffer pads each tensor to 16 bytes)."""
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(5, 7, 3, bias=True)
        self.fc = nn.Linear(13, 9)
        self.norm = nn.LayerNorm(9)
        self.gamma = nn.Parameter(torch.full((7,), 1e-6))
        self.head = nn.Conv2d(7, 1, 3, bias=False)


@pytest.mark.parametrize('world_scale', [1.0, 0.25])
def test_fused_adamw_matches_torch_adamw(world_scale):
    from slowtv_monodepth_b200.optim import FlatAdamW
    torch.manual_seed(0)
    ref = _Toy().double()
    ours = _Toy()
    ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    ours = ours.cuda()
    lr, wd = 3e-3, 1e-2
    decay = [p for n, p in ref.named_parameters() if not (p.ndim <= 1 or n.endswith('.bias'))]
    no_decay = [p for n, p in ref.named_parameters() if p.ndim <= 1 or n.endswith('.bias')]
    topt = torch.optim.AdamW([{'params': decay, 'weight_decay': wd}, {'params': no_decay, 'weight_decay': 0.}], lr=lr)
    opt = FlatAdamW(ours, lr=lr, weight_decay=wd)
    opt.world = round(1/world_scale)  # the kernel folds the 1/world gradient scale of the summed all-reduce
    names = [n for n, _ in ref.named_parameters()]
    g = torch.Generator().manual_seed(1)
    for step in range(5):
        opt.zero_grad()
        for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
            gr = torch.randn(pr.shape, generator=g, dtype=torch.float64)*(10.0**(step - 2))  # five decades of gradient scale
            pr.grad = gr.clone()
            po.grad.copy_((gr/world_scale).float())       # what a SUM all-reduce over `world` ranks would leave in the buffer
        topt.step()
        opt.step()
        torch.cuda.synchronize()
        for n, (pr, po) in zip(names, zip(ref.parameters(), ours.parameters())):
            err = (po.detach().double().cpu() - pr.detach()).abs().max().item()
            scale = pr.detach().abs().max().item() + lr
            assert err <= 1e-6*scale + 2e-7, f'step {step} {n}: |diff| {err:.3e} (scale {scale:.3e})'
    assert opt.step_count == 5
