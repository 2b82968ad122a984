"""The fused residual-network plan, run UNCHANGED on the CPU in fp32 through the reference ops
(ops/reference.py) and the emulated update (runtime/emulation.py), must reproduce autograd + torch.optim exactly:
this checks the orchestration (buffer wiring, residual gradient routing, strided dgrad, channels-last weight
layout, arena offsets, BatchNorm bookkeeping) independently of kernel numerics."""
import copy
import sys
import os

import pytest
import torch
import torch.nn as nn

sys.path.insert(1, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.nets import ResNet  # noqa: E402

from sparktorch_b200.ops import reference  # noqa: E402
from sparktorch_b200.runtime.emulation import EmulatedUpdate  # noqa: E402
from sparktorch_b200.runtime.fused_resnet import FusedResNet, lower_resnet  # noqa: E402
from sparktorch_b200.runtime.optim import parse_optimizer  # noqa: E402


def _plan(model, batch, spec, image, inference=False):
    low = lower_resnet(model, 3 * image * image)
    assert low is not None
    return FusedResNet(model, low, "ce", spec, 1, 0, torch.device("cpu"), batch, 0, 0, 8,
                       backend_factory=lambda numel, wb: EmulatedUpdate(numel, spec, wb), ops=reference,
                       act_dtype=torch.float32, inference=inference)


@pytest.mark.parametrize("epilogue_stats", ["0", "1"])
def test_resnet_plan_matches_autograd_and_sgd(monkeypatch, epilogue_stats):
    monkeypatch.setenv("SPARKTORCH_BN_EPILOGUE", epilogue_stats)
    torch.manual_seed(0)
    image, ncls, batch = 64, 16, 8
    model = ResNet(ncls, image)
    ref = copy.deepcopy(model).train()
    opt_kw = {"lr": 0.01, "momentum": 0.9, "weight_decay": 1e-4}
    spec = parse_optimizer(torch.optim.SGD, opt_kw)
    plan = _plan(model, batch, spec, image)
    plan.fu.updater.publish()
    ref_opt = torch.optim.SGD(ref.parameters(), **opt_kw)
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        x = torch.randn(batch, 3 * image * image, generator=g)
        y = torch.randint(0, ncls, (batch,), generator=g)
        plan.train["acts"][0].copy_(x)
        plan.train["tgt"].copy_(y.to(torch.int32))
        plan.forward(plan.train, slot=plan.loss_slot, train=True)
        plan.backward(plan.train)
        ref_opt.zero_grad()
        out = ref(x)
        loss = nn.functional.cross_entropy(out, y)
        loss.backward()
        # step 0 starts from identical weights: exact up to fp32 summation order.  Later steps compare two fp32
        # trajectories of a freshly initialised batch-norm net, which drift apart slowly (rounding is amplified).
        tight = step == 0
        assert abs(float(plan.loss_slot[0]) - float(loss.detach())) < (1e-4 if tight else 1e-2)
        torch.testing.assert_close(plan.train["logits"], out.detach(), rtol=1e-3, atol=1e-4 if tight else 2e-2)
        got = dict(model.named_parameters())
        for name, p in ref.named_parameters():
            a, b = got[name].grad.flatten().double(), p.grad.flatten().double()
            cos = float(a @ b / (a.norm() * b.norm() + 1e-30))
            assert cos > (0.999999 if tight else 0.995), (step, name, cos)
            assert abs(float(a.norm() / (b.norm() + 1e-30)) - 1.0) < (1e-4 if tight else 5e-2), (step, name)
        for (n1, b1), (_n2, b2) in zip(model.named_buffers(), ref.named_buffers()):
            if "running" in n1:
                torch.testing.assert_close(b1, b2, rtol=1e-3, atol=1e-4 if tight else 1e-2, msg=lambda m, n=n1: f"{n}: {m}")
        before = {n: p.detach().clone() for n, p in ref.named_parameters()}
        mine_before = {n: got[n].detach().clone() for n in before}
        plan.update()
        ref_opt.step()
        if tight:       # the update itself (momentum buffer seeding, weight decay, lr) on identical inputs
            for name, p in ref.named_parameters():
                torch.testing.assert_close(got[name].detach() - mine_before[name], p.detach() - before[name], rtol=1e-3, atol=1e-7,
                                           msg=lambda m, n=name: f"{n}: {m}")
    # the 3x3 weights live channels-last in the flat arena, the module still sees [Cout, Cin, 3, 3]
    w = model.layers[0].conv2.weight
    assert w.shape == (16, 16, 3, 3) and w.stride() == (144, 1, 48, 16)     # channels_last view of the arena slice
    assert plan.fu.steps == 3 and float(plan.fu.grad.abs().max()) == 0.0


def test_resnet_plan_inference_uses_running_statistics():
    torch.manual_seed(2)
    image, ncls, batch = 32, 10, 5
    model = ResNet(ncls, image)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.3)
    ref = copy.deepcopy(model).eval()
    spec = parse_optimizer(torch.optim.SGD, {"lr": 0.0})
    plan = _plan(model, batch, spec, image, inference=True)
    plan.fu.updater.publish()
    x = torch.randn(batch, 3 * image * image)
    plan.train["acts"][0].copy_(x)
    plan.forward(plan.train, slot=None, train=False)
    with torch.no_grad():
        want = ref(x)
    torch.testing.assert_close(plan.train["logits"], want, rtol=1e-3, atol=1e-4)
    for (n1, b1), (_n2, b2) in zip(model.named_buffers(), ref.named_buffers()):
        torch.testing.assert_close(b1, b2, rtol=0, atol=0)          # evaluation never touches the statistics
