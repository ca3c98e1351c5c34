"""CPU: the AdamW restatement in oracle/optim.py against the UNMODIFIED reference class (lib/helpers/optimizer_helper.py)."""
import os
import sys

import pytest
import torch

from oracle.optim import adamw_reference_step


@pytest.mark.reference
def test_adamw_port_matches_reference_class():
    sys.path.insert(0, "/root/reference")
    import warnings
    warnings.filterwarnings("ignore")
    from lib.helpers.optimizer_helper import build_optimizer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.LayerNorm(5), torch.nn.Linear(5, 3))
    ref_opt = build_optimizer({"type": "adamw", "lr": 2e-4, "weight_decay": 1e-4}, model)
    mine = [p.detach().clone() for p in model.parameters()]
    names = [n for n, _ in model.named_parameters()]
    ms = [torch.zeros_like(p) for p in mine]
    vs = [torch.zeros_like(p) for p in mine]
    wds = [0.0 if "bias" in n else 1e-4 for n in names]
    for step in range(1, 6):
        for p in model.parameters():
            p.grad = torch.randn_like(p)
        grads = [p.grad.clone() for p in model.parameters()]
        try:
            ref_opt.step()
        except TypeError:
            pytest.skip("the reference's deprecated add_(Number, Tensor) overloads are rejected by this torch build")
        adamw_reference_step(mine, grads, ms, vs, step, 2e-4, 0.9, 0.999, 1e-8, wds)
        for a, b in zip(model.parameters(), mine):
            assert torch.equal(a.detach(), b)
