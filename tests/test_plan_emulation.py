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


# ----------------------------------------------------------------------------------------------------------------
# MLP and CNN plans
from sparktorch_b200.runtime.fused_cnn import FusedCNN, lower_cnn  # noqa: E402
from sparktorch_b200.runtime.fused_mlp import FusedMLP, lower_mlp, loss_kind  # noqa: E402


def _emulated(plan_cls, model, low, kind, spec, batch, **kw):
    return plan_cls(model, low, kind, spec, 1, 0, torch.device("cpu"), batch, 0, 0, 16,
                    backend_factory=lambda numel, wb: EmulatedUpdate(numel, spec, wb), ops=reference, act_dtype=torch.float32, **kw)


def _compare_steps(plan, model, ref, ref_opt, crit, batches, windowed=False, steps_tight=3):
    got = dict(model.named_parameters())
    for step, (x, y) in enumerate(batches):
        if not windowed:
            plan.train["acts"][0].copy_(x)
            plan.train["tgt"].copy_(y if plan.kind != "ce" else y.to(torch.int32))
        plan.forward(plan.train, slot=plan.loss_slot, train=True, windowed=windowed)
        plan.backward(plan.train, windowed=windowed)
        ref_opt.zero_grad()
        out = ref(x)
        loss = crit(out, y)
        loss.backward()
        assert abs(float(plan.loss_slot[0]) - float(loss.detach())) < 1e-4 * max(1.0, abs(float(loss.detach()))), (step, float(plan.loss_slot[0]), float(loss))
        for name, p in ref.named_parameters():
            torch.testing.assert_close(got[name].grad, p.grad, rtol=1e-3, atol=1e-5, msg=lambda m, n=name, s=step: f"step {s} {n}: {m}")
        plan.update()
        ref_opt.step()
        for name, p in ref.named_parameters():
            torch.testing.assert_close(got[name].detach(), p.detach(), rtol=1e-3, atol=1e-5, msg=lambda m, n=name, s=step: f"step {s} {n}: {m}")


@pytest.mark.parametrize("fuse_loss", ["1", "0"])
def test_mlp_plan_cross_entropy_adam(monkeypatch, fuse_loss):
    monkeypatch.setenv("SPARKTORCH_FUSE_LOSS", fuse_loss)
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(20, 24), nn.ReLU(), nn.Linear(24, 24), nn.ReLU(), nn.Linear(24, 5))
    ref = copy.deepcopy(model)
    kw = {"lr": 1e-2, "weight_decay": 1e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_mlp(model)
    plan = _emulated(FusedMLP, model, low, loss_kind(nn.CrossEntropyLoss(), low), spec, 12)
    assert plan.fuse_loss == (fuse_loss == "1")
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(12, 20, generator=g), torch.randint(0, 5, (12,), generator=g)) for _ in range(4)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches)


def test_mlp_plan_mse_sgd_momentum_no_bias():
    torch.manual_seed(1)
    model = nn.Sequential(nn.Linear(10, 16, bias=False), nn.ReLU(), nn.Linear(16, 3))
    ref = copy.deepcopy(model)
    kw = {"lr": 0.05, "momentum": 0.9, "nesterov": True}
    spec = parse_optimizer(torch.optim.SGD, kw)
    low = lower_mlp(model)
    plan = _emulated(FusedMLP, model, low, loss_kind(nn.MSELoss(), low), spec, 9)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(4)
    batches = [(torch.randn(9, 10, generator=g), torch.randn(9, 3, generator=g)) for _ in range(4)]
    _compare_steps(plan, model, ref, torch.optim.SGD(ref.parameters(), **kw), nn.MSELoss(), batches)


def test_mlp_plan_batch_windows_follow_the_device_iteration_counter():
    """'sweep' sampling: the GEMMs address window (iter % n_windows) of the resident partition themselves."""
    torch.manual_seed(2)
    model = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4))
    ref = copy.deepcopy(model)
    kw = {"lr": 1e-2}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_mlp(model)
    B, windows = 6, 3
    plan = _emulated(FusedMLP, model, low, "ce", spec, B)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(5)
    X, Y = torch.randn(B * windows, 8, generator=g), torch.randint(0, 4, (B * windows,), generator=g)
    plan.train["acts"][0] = X.clone()                     # the resident partition IS the first operand
    plan.train["tgt"] = Y.to(torch.int32)
    plan.window = (windows, B)
    batches = [(X[(i % windows) * B: (i % windows + 1) * B], Y[(i % windows) * B: (i % windows + 1) * B]) for i in range(5)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches, windowed=True)


class _ConvNet(nn.Module):
    """conv-relu-pool, conv-relu (no pool), conv-relu-pool, flatten, two Linear layers."""

    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(1, 8, 3, padding=1)
        self.c2 = nn.Conv2d(8, 16, 3)
        self.c3 = nn.Conv2d(16, 16, 3, padding=1)
        self.f1 = nn.Linear(16 * 2 * 2, 20)
        self.f2 = nn.Linear(20, 6)

    def forward(self, x):
        x = x.view(-1, 1, 12, 12)
        x = nn.functional.max_pool2d(torch.relu(self.c1(x)), 2)      # 8 x 6 x 6
        x = torch.relu(self.c2(x))                                   # 16 x 4 x 4
        x = nn.functional.max_pool2d(torch.relu(self.c3(x)), 2)      # 16 x 2 x 2
        x = torch.flatten(x, 1)
        return self.f2(torch.relu(self.f1(x)))


@pytest.mark.parametrize("fuse_loss", ["1", "0"])
def test_cnn_plan_matches_autograd(monkeypatch, fuse_loss):
    monkeypatch.setenv("SPARKTORCH_FUSE_LOSS", fuse_loss)
    torch.manual_seed(3)
    model = _ConvNet()
    ref = copy.deepcopy(model)
    kw = {"lr": 5e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_cnn(model, 144)
    assert low is not None and [c.pool for c in low.convs] == [2, 0, 2]
    plan = _emulated(FusedCNN, model, low, "ce", spec, 7)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(6)
    batches = [(torch.randn(7, 144, generator=g), torch.randint(0, 6, (7,), generator=g)) for _ in range(3)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches)


def test_mnist_cnn_plan_without_dropout_matches_autograd():
    from sparktorch_b200.models import MnistCNN

    torch.manual_seed(4)
    model = MnistCNN()
    for m in model.modules():
        if isinstance(m, (nn.Dropout, nn.Dropout2d)):
            m.p = 0.0                                                # the counter-based dropout mask is not emulated
    ref = copy.deepcopy(model)
    kw = {"lr": 1e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_cnn(model, 784)
    assert low is not None
    plan = _emulated(FusedCNN, model, low, "ce", spec, 4)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(7)
    batches = [(torch.randn(4, 784, generator=g), torch.randint(0, 10, (4,), generator=g)) for _ in range(2)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches)


@pytest.mark.parametrize("backend", ["sync", "hogwild"])
def test_cnn_plan_split_head_reads_labels_of_the_current_window(backend):
    """MnistCNN's 3872-wide head takes the split-K + stand-alone loss path; with 'sweep' windows the loss kernel must read
    the labels of the SAME window the im2col reads the images from (both for the synchronous plan and for the hogwild
    worker, whose forward GEMMs additionally gather the weights from their owners)."""
    from sparktorch_b200.models import MnistCNN
    from sparktorch_b200.runtime.emulation import EmulatedHogwild

    torch.manual_seed(5)
    model = MnistCNN(p_drop=0.0)
    ref = copy.deepcopy(model)
    kw = {"lr": 1e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_cnn(model, 784)
    B, windows = 4, 3
    if backend == "sync":
        plan = _emulated(FusedCNN, model, low, "ce", spec, B)
    else:
        plan = FusedCNN(model, low, "ce", spec, 1, 0, torch.device("cpu"), B, 0, 0, 16,
                        backend_factory=lambda n, w: EmulatedHogwild(n, spec, w, world=2), ops=reference, act_dtype=torch.float32)
    assert plan._deep_head(B)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(8)
    X, Y = torch.randn(B * windows, 784, generator=g), torch.randint(0, 10, (B * windows,), generator=g)
    plan.train["acts"][0] = X.clone()
    plan.train["tgt"] = Y.to(torch.int32)
    plan.window = (windows, B)
    batches = [(X[(i % windows) * B: (i % windows + 1) * B], Y[(i % windows) * B: (i % windows + 1) * B]) for i in range(4)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches, windowed=True)


def test_cnn_plan_fused_loss_with_batch_windows():
    torch.manual_seed(6)
    model = _ConvNet()
    ref = copy.deepcopy(model)
    kw = {"lr": 5e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_cnn(model, 144)
    B, windows = 5, 4
    plan = _emulated(FusedCNN, model, low, "ce", spec, B)
    assert plan.fuse_loss and not plan._deep_head(B)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(9)
    X, Y = torch.randn(B * windows, 144, generator=g), torch.randint(0, 6, (B * windows,), generator=g)
    plan.train["acts"][0] = X.clone()
    plan.train["tgt"] = Y.to(torch.int32)
    plan.window = (windows, B)
    batches = [(X[(i % windows) * B: (i % windows + 1) * B], Y[(i % windows) * B: (i % windows + 1) * B]) for i in range(6)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches, windowed=True)


def test_mnist_cnn_plan_with_counter_based_dropout2d():
    """Dropout2d in the fused plan is a pure function of (seed, iteration, sample, channel).  Applying the same keep
    mask to an autograd copy of the network must give the same loss, gradients and updates."""
    from sparktorch_b200.models import MnistCNN

    class MaskedDrop(nn.Module):
        def __init__(self):
            super().__init__()
            self.mask = None

        def forward(self, x):
            return x * self.mask

    torch.manual_seed(7)
    model = MnistCNN(p_drop=0.25)
    ref = copy.deepcopy(model)
    ref.drop = MaskedDrop()
    kw = {"lr": 1e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    low = lower_cnn(model, 784)
    assert low.convs[-1].p_drop == 0.25
    B, seed = 6, 11
    plan = _emulated(FusedCNN, model, low, "ce", spec, B, seed=seed)
    plan.fu.updater.publish()
    ref_opt = torch.optim.Adam(ref.parameters(), **kw)
    g = torch.Generator().manual_seed(10)
    got = dict(model.named_parameters())
    kept = []
    for step in range(3):
        x, y = torch.randn(B, 784, generator=g), torch.randint(0, 10, (B,), generator=g)
        keep = reference.dropout2d_keep_mask(seed + 1000003 * (len(low.convs) - 1), step, B, 32, 0.25)
        kept.append(float(keep.float().mean()))
        ref.drop.mask = (keep.float() / 0.75).view(B, 32, 1, 1)
        plan.train["acts"][0].copy_(x)
        plan.train["tgt"].copy_(y.to(torch.int32))
        plan.forward(plan.train, slot=plan.loss_slot, train=True)
        plan.backward(plan.train)
        ref_opt.zero_grad()
        loss = nn.functional.cross_entropy(ref(x), y)
        loss.backward()
        assert abs(float(plan.loss_slot[0]) - float(loss.detach())) < 1e-4
        for name, p in ref.named_parameters():
            torch.testing.assert_close(got[name].grad, p.grad, rtol=1e-3, atol=1e-5, msg=lambda m, n=name: f"{n}: {m}")
        plan.update()
        ref_opt.step()
    assert 0.5 < sum(kept) / len(kept) < 0.95 and len(set(kept)) > 1        # masks vary per iteration, ~75 % kept


def test_mlp_plan_randomised_architectures_and_optimizers():
    """30 random MLPs (ragged widths, optional biases / ReLUs, CE or MSE, SGD / Adam / AdamW variants, with and without
    batch windows) - the plan must track autograd + torch.optim step by step."""
    import random

    rnd = random.Random(1234)
    for case in range(30):
        torch.manual_seed(100 + case)
        depth = rnd.randint(1, 4)
        dims = [rnd.choice([3, 8, 10, 17, 24, 33])] + [rnd.choice([5, 8, 12, 20, 31]) for _ in range(depth - 1)]
        kind = rnd.choice(["ce", "mse"])
        out_dim = rnd.choice([2, 5, 10, 32, 40]) if kind == "ce" else rnd.choice([1, 3, 8])
        dims.append(out_dim)
        layers = []
        for i in range(depth):
            layers.append(nn.Linear(dims[i], dims[i + 1], bias=rnd.random() < 0.8))
            if i < depth - 1 and rnd.random() < 0.8:
                layers.append(nn.ReLU())
        model = nn.Sequential(*layers)
        ref = copy.deepcopy(model)
        opt = rnd.choice(["sgd", "sgd_m", "sgd_nest", "adam", "adam_wd", "adamw"])
        if opt == "sgd":
            cls, kw = torch.optim.SGD, {"lr": 0.05, "weight_decay": rnd.choice([0.0, 1e-3])}
        elif opt == "sgd_m":
            cls, kw = torch.optim.SGD, {"lr": 0.03, "momentum": 0.9, "dampening": rnd.choice([0.0, 0.1])}
        elif opt == "sgd_nest":
            cls, kw = torch.optim.SGD, {"lr": 0.03, "momentum": 0.8, "nesterov": True}
        elif opt == "adam":
            cls, kw = torch.optim.Adam, {"lr": 3e-3, "betas": (0.8, 0.95)}
        elif opt == "adam_wd":
            cls, kw = torch.optim.Adam, {"lr": 3e-3, "weight_decay": 1e-2}
        else:
            cls, kw = torch.optim.AdamW, {"lr": 3e-3, "weight_decay": 5e-2}
        spec = parse_optimizer(cls, kw)
        low = lower_mlp(model)
        assert low is not None, case
        crit = nn.CrossEntropyLoss() if kind == "ce" else nn.MSELoss()
        B = rnd.choice([4, 9, 16])
        windowed = rnd.random() < 0.5
        plan = _emulated(FusedMLP, model, low, loss_kind(crit, low), spec, B)
        plan.fu.updater.publish()
        windowed = windowed and plan.fuse_loss              # windows need the loss fused in the last GEMM
        g = torch.Generator().manual_seed(500 + case)
        n_win = 3
        X = torch.randn(B * n_win, dims[0], generator=g)
        Y = torch.randint(0, out_dim, (B * n_win,), generator=g) if kind == "ce" else torch.randn(B * n_win, out_dim, generator=g)
        batches = [(X[(i % n_win) * B: (i % n_win + 1) * B], Y[(i % n_win) * B: (i % n_win + 1) * B]) for i in range(4)]
        if windowed:
            plan.train["acts"][0] = X.clone()
            plan.train["tgt"] = Y.to(torch.int32) if kind == "ce" else Y.clone()
            plan.window = (n_win, B)
        try:
            _compare_steps(plan, model, ref, cls(ref.parameters(), **kw), crit, batches, windowed=windowed)
        except AssertionError as e:
            raise AssertionError(f"case {case}: dims {dims} kind {kind} opt {opt} {kw} B {B} windowed {windowed}: {e}") from None


class _RandCNN(nn.Module):
    def __init__(self, in_shape, conv_specs, head):
        super().__init__()
        self.in_shape = in_shape
        c = in_shape[0]
        self.convs = nn.ModuleList()
        self.meta = []
        for (cout, k, s, p, relu, pool) in conv_specs:
            self.convs.append(nn.Conv2d(c, cout, k, stride=s, padding=p, bias=True))
            self.meta.append((relu, pool))
            c = cout
        self.head = nn.ModuleList([nn.Linear(a, b) for a, b in head])

    def forward(self, x):
        x = x.view(-1, self.in_shape[0], self.in_shape[1], self.in_shape[2])
        for conv, (relu, pool) in zip(self.convs, self.meta):
            x = conv(x)
            if relu:
                x = torch.relu(x)
            if pool:
                x = nn.functional.max_pool2d(x, pool)
        x = torch.flatten(x, 1)
        for i, lin in enumerate(self.head):
            x = lin(x)
            if i < len(self.head) - 1:
                x = torch.relu(x)
        return x


def test_cnn_plan_randomised_geometries():
    import random

    rnd = random.Random(99)
    done = 0
    for case in range(40):
        C0, H0 = rnd.choice([1, 3]), rnd.choice([10, 12, 14])
        n_conv = rnd.randint(1, 3)
        specs, c, h = [], C0, H0
        ok = True
        for i in range(n_conv):
            k = rnd.choice([3, 5])
            s = rnd.choice([1, 2]) if i == 0 else 1
            p = rnd.choice([0, 1, 2])
            cout = rnd.choice([4, 8, 12])
            pool = rnd.choice([0, 2])
            oh = (h + 2 * p - k) // s + 1
            if oh < 2:
                ok = False
                break
            if pool:
                oh //= pool
            if oh < 1:
                ok = False
                break
            specs.append((cout, k, s, p, rnd.random() < 0.8, pool))
            c, h = cout, oh
        if not ok:
            continue
        feat = c * h * h
        ncls = rnd.choice([3, 6, 10])
        head = [(feat, ncls)] if rnd.random() < 0.5 else [(feat, 16), (16, ncls)]
        torch.manual_seed(300 + case)
        model = _RandCNN((C0, H0, H0), specs, head)
        low = lower_cnn(model, C0 * H0 * H0)
        if low is None:
            continue
        ref = copy.deepcopy(model)
        kw = {"lr": 3e-3}
        spec = parse_optimizer(torch.optim.Adam, kw)
        B = rnd.choice([3, 6])
        plan = _emulated(FusedCNN, model, low, "ce", spec, B)
        plan.fu.updater.publish()
        g = torch.Generator().manual_seed(700 + case)
        batches = [(torch.randn(B, C0 * H0 * H0, generator=g), torch.randint(0, ncls, (B,), generator=g)) for _ in range(2)]
        try:
            _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches)
        except AssertionError as e:
            raise AssertionError(f"case {case}: in {(C0, H0)} convs {specs} head {head} B {B}: {e}") from None
        done += 1
    assert done >= 15, done


@pytest.mark.parametrize("which", ["mlp", "cnn", "resnet"])
def test_validation_forward_fills_the_shared_loss_slot_and_leaves_gradients_alone(which):
    """With validationPct the step is train fwd (loss -> tail[2]) + bwd + validation fwd (loss -> tail[0], the slot that
    rides in the all-reduce and drives early stopping) + update."""
    torch.manual_seed(8)
    spec = parse_optimizer(torch.optim.Adam, {"lr": 1e-3})
    B, NV = 6, 10
    if which == "mlp":
        model = nn.Sequential(nn.Linear(12, 16), nn.ReLU(), nn.Linear(16, 4))
        low, d0, ncls = lower_mlp(model), 12, 4
        plan = FusedMLP(model, low, "ce", spec, 1, 0, torch.device("cpu"), B, NV, 0, 8,
                        backend_factory=lambda n, w: EmulatedUpdate(n, spec, w), ops=reference, act_dtype=torch.float32)
    elif which == "cnn":
        model = _ConvNet()
        low, d0, ncls = lower_cnn(model, 144), 144, 6
        plan = FusedCNN(model, low, "ce", spec, 1, 0, torch.device("cpu"), B, NV, 0, 8,
                        backend_factory=lambda n, w: EmulatedUpdate(n, spec, w), ops=reference, act_dtype=torch.float32)
    else:
        model = ResNet(8, 32)
        low, d0, ncls = lower_resnet(model, 3 * 32 * 32), 3 * 32 * 32, 8
        plan = FusedResNet(model, low, "ce", spec, 1, 0, torch.device("cpu"), B, NV, 0, 8,
                           backend_factory=lambda n, w: EmulatedUpdate(n, spec, w), ops=reference, act_dtype=torch.float32)
    ref = copy.deepcopy(model)
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(12)
    x, y = torch.randn(B, d0, generator=g), torch.randint(0, ncls, (B,), generator=g)
    xv, yv = torch.randn(NV, d0, generator=g), torch.randint(0, ncls, (NV,), generator=g)
    plan.train["acts"][0].copy_(x)
    plan.train["tgt"].copy_(y.to(torch.int32))
    plan.val["acts"][0].copy_(xv)
    plan.val["tgt"].copy_(yv.to(torch.int32))
    plan.forward(plan.train, slot=plan.train_loss_slot, train=True)
    plan.backward(plan.train)
    grads = plan.fu.grad[: plan.fu.numel].clone()
    bn_before = {n: b.clone() for n, b in model.named_buffers()}
    plan.forward(plan.val, slot=plan.loss_slot, train=False)
    torch.testing.assert_close(plan.fu.grad[: plan.fu.numel], grads, rtol=0, atol=0)        # validation adds no gradient
    for n, b in model.named_buffers():
        torch.testing.assert_close(b, bn_before[n], rtol=0, atol=0)                           # ... and no BN statistics
    loss_train = nn.functional.cross_entropy(ref(x), y)
    with torch.no_grad():
        loss_val = nn.functional.cross_entropy(ref(xv), yv)      # train-mode BN on the validation batch, like the plan
    assert abs(float(plan.train_loss_slot[0]) - float(loss_train.detach())) < 1e-4
    assert abs(float(plan.loss_slot[0]) - float(loss_val)) < 1e-4


@pytest.mark.parametrize("which", ["mlp", "cnn"])
def test_k_ag_forward_gathers_weights_from_their_owners(which):
    """Hogwild layout: parameters sharded over W owners (every other copy poisoned with NaN); the forward GEMMs of the
    plan gather weights AND biases through peer_table, the dgrad GEMMs use the gathered local copy."""
    from sparktorch_b200.runtime.emulation import EmulatedHogwild

    torch.manual_seed(9)
    kw = {"lr": 2e-3}
    spec = parse_optimizer(torch.optim.Adam, kw)
    B = 5
    if which == "mlp":
        model = nn.Sequential(nn.Linear(12, 20), nn.ReLU(), nn.Linear(20, 16), nn.ReLU(), nn.Linear(16, 4))
        low, d0, ncls, cls = lower_mlp(model), 12, 4, FusedMLP
    else:
        model = _ConvNet()
        low, d0, ncls, cls = lower_cnn(model, 144), 144, 6, FusedCNN
    ref = copy.deepcopy(model)
    plan = cls(model, low, "ce", spec, 1, 0, torch.device("cpu"), B, 0, 0, 16,
               backend_factory=lambda n, w: EmulatedHogwild(n, spec, w, world=3), ops=reference, act_dtype=torch.float32)
    assert all(plan.peer_kw[k] for k in (plan.peer_kw if isinstance(plan.peer_kw, dict) else range(len(plan.peer_kw))))
    plan.fu.updater.publish()
    g = torch.Generator().manual_seed(13)
    batches = [(torch.randn(B, d0, generator=g), torch.randint(0, ncls, (B,), generator=g)) for _ in range(3)]
    _compare_steps(plan, model, ref, torch.optim.Adam(ref.parameters(), **kw), nn.CrossEntropyLoss(), batches)
    assert not torch.isnan(plan.fu.param).any()
