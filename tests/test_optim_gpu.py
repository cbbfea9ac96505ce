"""csrc/optim.cu through parallel.FlatAdamW against torch.optim.AdamW (the optimizer the reference builds, optimizer.py:4-26)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.LayerNorm(64), torch.nn.ReLU(), torch.nn.Linear(64, 19),
                               torch.nn.Linear(19, 3, bias=False)).cuda()


def test_flat_adamw_matches_torch_adamw_with_the_reference_param_groups():
    from vdetr_b200 import parallel
    no_decay = lambda n, p: p.dim() == 1 or n.endswith("bias")          # noqa: E731  (--filter_biases_wd, optimizer.py:11)
    a, b = _model(0), _model(0)
    ref = torch.optim.AdamW([{"params": [p for n, p in a.named_parameters() if no_decay(n, p)], "weight_decay": 0.0},
                             {"params": [p for n, p in a.named_parameters() if not no_decay(n, p)], "weight_decay": 0.1}], lr=7e-4)
    opt = parallel.FlatAdamW(b.named_parameters(), lr=7e-4, weight_decay=0.1, no_decay=no_decay)
    assert opt.n == sum(p.numel() for p in b.parameters()) and 0 < opt.n_decay < opt.n
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(5):
        x = torch.randn(50, 37, device="cuda", generator=g)
        if it == 3:                                                      # a learning-rate schedule step (engine.py:28-52)
            for grp in ref.param_groups:
                grp["lr"] = 3e-4
            opt.set_lr(3e-4)
        ref.zero_grad(set_to_none=True)
        a(x).square().sum().backward()
        ref.step()
        opt.zero_grad()
        b(x).square().sum().backward()
        opt.step()
        for (n, p), q in zip(a.named_parameters(), b.parameters()):
            assert (p - q).abs().max().item() <= 2e-6 * (p.abs().max().item() + 1e-3), (it, n)
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    assert float(sd["step"]) == 5.0


def test_flat_adamw_gradient_clipping_and_world_scale():
    """clip_grad_norm_ (engine.py:105-106) and the 1 / world factor of a SUM all-reduce are one device-side gradient scale."""
    from vdetr_b200 import parallel
    a, b = _model(2), _model(2)
    ref = torch.optim.AdamW(a.parameters(), lr=1e-3, weight_decay=0.05)
    opt = parallel.FlatAdamW(b.named_parameters(), lr=1e-3, weight_decay=0.05)
    opt.world_scale = 0.5                                                # as if two ranks had summed identical gradients
    x = torch.randn(40, 37, device="cuda")
    a(x).square().sum().backward()
    want_norm = torch.nn.utils.clip_grad_norm_(a.parameters(), 0.1)
    ref.step()
    opt.zero_grad()
    (2.0 * b(x).square().sum()).backward()                               # the "summed" gradient of two ranks
    got_norm = opt.clip_grad_norm_(0.1)
    opt.step()
    assert abs(float(got_norm) - float(want_norm)) <= 1e-4 * float(want_norm)
    for p, q in zip(a.parameters(), b.parameters()):
        assert (p - q).abs().max().item() <= 2e-6 * (p.abs().max().item() + 1e-3)


def test_flat_adamw_rejects_cpu_parameters():
    from vdetr_b200 import parallel
    with pytest.raises(RuntimeError):
        parallel.FlatAdamW(torch.nn.Linear(3, 3).named_parameters())
