"""Flat fused-update math vs torch.optim (the CUDA kernel is checked against the
same oracle in tests/test_gpu_kernels.py)."""
import dill
import pytest
import torch

from sparktorch_b200.runtime.optim import apply_update_, new_state, parse_optimizer

CASES = [
    (torch.optim.Adam, dict(lr=1e-2)),
    (torch.optim.Adam, dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)),
    (torch.optim.AdamW, dict(lr=1e-2, weight_decay=0.05)),
    (torch.optim.SGD, dict(lr=0.1)),
    (torch.optim.SGD, dict(lr=0.1, momentum=0.9)),
    (torch.optim.SGD, dict(lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.01)),
    (torch.optim.SGD, dict(lr=0.1, momentum=0.5, dampening=0.3)),
    (torch.optim.Adam, dict(lr=1e-2, maximize=True)),
]


@pytest.mark.parametrize("cls,kw", CASES)
def test_flat_update_matches_torch(cls, kw):
    torch.manual_seed(0)
    p_ref = torch.nn.Parameter(torch.randn(257))
    opt = cls([p_ref], **kw)
    spec = parse_optimizer(cls, kw)
    assert spec.kind != "other"
    p = p_ref.detach().clone()
    st = new_state(spec, p)
    for t in range(1, 8):
        g = torch.randn(257)
        p_ref.grad = g.clone()
        opt.step()
        apply_update_(spec, p, g, st, t)
        torch.testing.assert_close(p, p_ref.detach(), rtol=2e-5, atol=2e-6)


def test_parse_handles_dill_recreated_classes_and_fallbacks():
    # dill re-creates torch optimizer classes by value → identity checks fail
    clone = dill.loads(dill.dumps(torch.optim.Adam))
    assert parse_optimizer(clone, {"lr": 0.5}).kind == "adam"
    assert parse_optimizer(dill.loads(dill.dumps(torch.optim.SGD)), {"lr": 0.5, "momentum": 0.9}).n_state == 1
    assert parse_optimizer(torch.optim.RMSprop, {}).kind == "other"
    assert parse_optimizer(torch.optim.Adam, {"amsgrad": True}).kind == "other"

    class MyAdam(torch.optim.Adam):
        def step(self, closure=None):
            return super().step(closure)

    class PlainSub(torch.optim.Adam):
        pass

    assert parse_optimizer(MyAdam, {}).kind == "other"
    assert parse_optimizer(PlainSub, {}).kind == "adam"
