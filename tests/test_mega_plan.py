"""K-STEP task graph (runtime/mega_mlp.py) on the CPU: the builder's tables for W emulated replicas are interpreted by
``MegaEmulator`` under random interleavings and must reproduce autograd + torch.optim on the global batch - tile
coverage, batch windows, per-m-tile partial slots, shard ownership, flag masks, early stopping, dead-lock freedom and
cross-step buffer reuse are all exercised without a GPU (the CUDA kernel executes the same tables)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from sparktorch_b200.ops import reference
from sparktorch_b200.runtime.emulation import EmulatedUpdate
from sparktorch_b200.runtime.fused_mlp import FusedMLP, loss_kind, lower_mlp
from sparktorch_b200.runtime.mega_mlp import MS_MAX_FLAGS, MS_TASK, MegaEmulator, MegaPlan, mega_extra_elems
from sparktorch_b200.runtime.optim import parse_optimizer


def _ranks(model, crit, opt_cls, kw, W, B, patience=0, window=None, data=None, grid=148, chunk=0):
    """W emulated replicas of ``model``; data[r] = (X_r, Y_r) resident partition (or one batch) of rank r."""
    spec = parse_optimizer(opt_cls, kw)
    plans, megas = [], []
    for r in range(W):
        m = copy.deepcopy(model)
        low = lower_mlp(m)
        kind = loss_kind(crit, low)
        d = [low.layers[0].module.in_features] + [ls.module.out_features for ls in low.layers]
        extra = mega_extra_elems(d, [ls.module.bias is not None for ls in low.layers], B)
        plan = FusedMLP(m, low, kind, spec, W, r, torch.device("cpu"), B, 0, patience, 64,
                        backend_factory=lambda numel, wb: EmulatedUpdate(numel, spec, wb, grad_extra=extra),
                        ops=reference, act_dtype=torch.float32)
        plan.fu.updater.publish()
        X, Y = data[r]
        tgt = Y.to(torch.int32) if kind == "ce" else Y.to(torch.float32)
        numel = plan.fu.numel
        per = ((numel // 4 + W - 1) // W) * 4
        owned = (min(numel, per * r), min(numel, per * (r + 1)))
        flags = torch.zeros(4096, dtype=torch.int64)        # AR_FLAG_WORDS: epoch flags first, cross-GPU counters from word 1024
        mp = MegaPlan(plan, spec, patience, 0, owned, window, X.clone(), tgt, flags, grid=grid, chunk=chunk)
        assert mp.ok, mp.reason
        plans.append(plan)
        megas.append(mp)
    return plans, megas, spec


def _reference_run(model, crit, opt_cls, kw, batches):
    ref = copy.deepcopy(model)
    opt = opt_cls(ref.parameters(), **kw)
    losses = []
    for x, y in batches:
        opt.zero_grad()
        loss = crit(ref(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return ref, losses


def _check(plans, ref, rtol=2e-3, atol=2e-5):
    for plan in plans:
        got = dict(plan.arena.module.named_parameters())
        for name, p in ref.named_parameters():
            torch.testing.assert_close(got[name].detach(), p.detach(), rtol=rtol, atol=atol, msg=lambda m, n=name: f"{n}: {m}")


def test_task_struct_layout():
    assert MS_TASK.itemsize == 256
    assert MS_TASK.fields["bias"][1] == 64 and MS_TASK.fields["wait"][1] == 184 and MS_TASK.fields["sig"][1] == 232


@pytest.mark.parametrize("W", [1, 2, 3])
def test_mnist_like_mlp_ce_adam_full_batch(W):
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(40, 24), nn.ReLU(), nn.Linear(24, 24), nn.ReLU(), nn.Linear(24, 5))
    B, steps = 12, 4
    g = torch.Generator().manual_seed(1)
    data = [(torch.randn(B, 40, generator=g), torch.randint(0, 5, (B,), generator=g)) for _ in range(W)]
    kw = {"lr": 1e-2, "weight_decay": 1e-3}
    plans, megas, _ = _ranks(model, nn.CrossEntropyLoss(), torch.optim.Adam, kw, W, B, data=data, grid=7)
    emu = MegaEmulator(megas)
    emu.run(steps, seed=W)
    gx, gy = torch.cat([x for x, _ in data]), torch.cat([y for _, y in data])
    ref, losses = _reference_run(model, nn.CrossEntropyLoss(), torch.optim.Adam, kw, [(gx, gy)] * steps)
    _check(plans, ref)
    for st in emu.state:
        np.testing.assert_allclose(st["losses"], losses, rtol=1e-4)
        assert st["iter"] == steps and st["step"] == steps


@pytest.mark.parametrize("W,seed", [(1, 0), (2, 1), (2, 2), (4, 3)])
def test_batch_windows_multiple_m_tiles_and_wide_layers(W, seed):
    """B = 300 rows -> 3 m-tiles (partial slots), 200-wide layers -> several n-tiles / wgrad tiles, sweep windows."""
    torch.manual_seed(seed)
    model = nn.Sequential(nn.Linear(70, 200), nn.ReLU(), nn.Linear(200, 130), nn.ReLU(), nn.Linear(130, 10))
    B, windows, steps = 300, 2, 3
    g = torch.Generator().manual_seed(10 + seed)
    data = [(torch.randn(B * windows, 70, generator=g), torch.randint(0, 10, (B * windows,), generator=g)) for _ in range(W)]
    kw = {"lr": 3e-3}
    plans, megas, _ = _ranks(model, nn.CrossEntropyLoss(), torch.optim.Adam, kw, W, B, window=(windows, B), data=data, grid=16)
    emu = MegaEmulator(megas)
    emu.run(steps, seed=seed)
    batches = []
    for i in range(steps):
        w = i % windows
        batches.append((torch.cat([x[w * B:(w + 1) * B] for x, _ in data]), torch.cat([y[w * B:(w + 1) * B] for _, y in data])))
    ref, losses = _reference_run(model, nn.CrossEntropyLoss(), torch.optim.Adam, kw, batches)
    _check(plans, ref)
    np.testing.assert_allclose(emu.state[0]["losses"], losses, rtol=1e-4)


def test_mse_sgd_momentum_no_bias_and_launch_split():
    """two launches (2 + 3 steps) continue each other: epochs / iterations / optimizer steps are launch parameters."""
    torch.manual_seed(3)
    model = nn.Sequential(nn.Linear(10, 16, bias=False), nn.ReLU(), nn.Linear(16, 3))
    B, W = 9, 2
    g = torch.Generator().manual_seed(4)
    data = [(torch.randn(B, 10, generator=g), torch.randn(B, 3, generator=g)) for _ in range(W)]
    kw = {"lr": 0.05, "momentum": 0.9, "nesterov": True}
    plans, megas, _ = _ranks(model, nn.MSELoss(), torch.optim.SGD, kw, W, B, data=data, grid=5)
    emu = MegaEmulator(megas)
    emu.run(2, iter0=0, step0=0, epoch0=0, seed=5)
    emu.run(3, iter0=2, step0=2, epoch0=2, seed=6)
    gx, gy = torch.cat([x for x, _ in data]), torch.cat([y for _, y in data])
    ref, losses = _reference_run(model, nn.MSELoss(), torch.optim.SGD, kw, [(gx, gy)] * 5)
    _check(plans, ref)
    np.testing.assert_allclose(emu.state[1]["losses"], losses, rtol=1e-4)


def test_early_stopping_skips_the_stopping_update_and_freezes():
    torch.manual_seed(5)
    model = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 3))
    B, W = 8, 2
    g = torch.Generator().manual_seed(6)
    data = [(torch.randn(B, 6, generator=g), torch.randint(0, 3, (B,), generator=g)) for _ in range(W)]
    kw = {"lr": 0.1, "maximize": True}  # gradient ASCENT: the loss rises, patience runs out
    plans, megas, _ = _ranks(model, nn.CrossEntropyLoss(), torch.optim.SGD, kw, W, B, patience=2, data=data, grid=4)
    emu = MegaEmulator(megas)
    emu.run(12, seed=7)
    # oracle: reference semantics (distributed.py:185-198): the stopping iteration's update is skipped
    from sparktorch_b200.early_stopper import EarlyStopping

    ref = copy.deepcopy(model)
    opt = torch.optim.SGD(ref.parameters(), **kw)
    es = EarlyStopping(patience=2)
    gx, gy = torch.cat([x for x, _ in data]), torch.cat([y for _, y in data])
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = nn.CrossEntropyLoss()(ref(gx), gy)
        loss.backward()
        losses.append(float(loss.detach()))
        if es.step(float(loss.detach())):
            break
        opt.step()
    assert len(losses) < 12, "test setup: early stopping must trigger"
    for st in emu.state:
        assert st["stopped"] and st["iter"] == len(losses) and st["step"] == len(losses) - 1
        np.testing.assert_allclose(st["losses"], losses, rtol=1e-4)
    _check(plans, ref)


def test_randomised_architectures():
    rng = np.random.default_rng(0)
    for trial in range(6):
        torch.manual_seed(100 + trial)
        L = int(rng.integers(1, 5))
        dims = [int(rng.integers(3, 90))] + [int(rng.integers(4, 150)) for _ in range(L - 1)] + [int(rng.integers(2, 30))]
        layers = []
        for i in range(L):
            layers.append(nn.Linear(dims[i], dims[i + 1], bias=bool(rng.integers(0, 2))))
            if i < L - 1 and rng.integers(0, 4) > 0:
                layers.append(nn.ReLU())
        model = nn.Sequential(*layers)
        W = int(rng.integers(1, 4))
        B = int(rng.integers(5, 200))
        ce = bool(rng.integers(0, 2))
        crit = nn.CrossEntropyLoss() if ce else nn.MSELoss()
        g = torch.Generator().manual_seed(trial)
        data = [(torch.randn(B, dims[0], generator=g),
                 torch.randint(0, dims[-1], (B,), generator=g) if ce else torch.randn(B, dims[-1], generator=g)) for _ in range(W)]
        # (eps = 1e-4: with the default 1e-8 Adam turns fp32 rounding noise on near-zero gradients into lr-sized differences)
        opt_cls, kw = [(torch.optim.Adam, {"lr": 1e-2, "eps": 1e-4}), (torch.optim.AdamW, {"lr": 1e-2, "weight_decay": 0.1, "eps": 1e-4}),
                       (torch.optim.SGD, {"lr": 0.05, "momentum": 0.8})][trial % 3]
        plans, megas, _ = _ranks(model, crit, opt_cls, kw, W, B, data=data, grid=int(rng.integers(3, 40)), chunk=512)
        emu = MegaEmulator(megas)
        emu.run(3, seed=trial)
        gx, gy = torch.cat([x for x, _ in data]), torch.cat([y for _, y in data])
        ref, losses = _reference_run(model, crit, opt_cls, kw, [(gx, gy)] * 3)
        _check(plans, ref)
        np.testing.assert_allclose(emu.state[0]["losses"], losses, rtol=2e-4)


def test_mnist_dnn_plan_shape():
    """the flagship configuration: task counts, placement and the 148-CTA grid"""
    from sparktorch_b200.models import mnist_dnn

    torch.manual_seed(0)
    model = mnist_dnn()
    B = 256
    data = [(torch.zeros(B * 2, 784), torch.zeros(B * 2, dtype=torch.long))]
    plans, megas, _ = _ranks(model, nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, 1, B, window=(2, B), data=data)
    info = megas[0].describe()
    # 26 wgrad tiles of 128 x 64, drained by two CTAs each (32-column windows; the last column tile has 16 valid columns)
    assert info["groups"]["F1"] == 16 and info["groups"]["H3"] == 16 and info["groups"]["F3"] == 2 and info["groups"]["G1"] == 50 and info["groups"]["D"] == 1
    assert info["groups"]["B2"] == 16 and info["groups"]["B3"] == 16
    assert info["ctas"] <= 148 and info["max_tasks_per_cta"] <= 24


@pytest.mark.parametrize("flags,fused", [("0", "1"), ("0", "0"), ("1", "1")])
def test_single_gpu_variants_and_early_stop(monkeypatch, flags, fused):
    """one GPU: (a) optimizer fused into the weight-gradient tiles, (b) separate update chunks on counters, (c) the flag
    protocol of the multi-GPU case with one rank - all with early stopping (the stopping step's update is skipped)."""
    monkeypatch.setenv("SPARKTORCH_MEGA_FLAGS", flags)
    monkeypatch.setenv("SPARKTORCH_MEGA_FUSED_OPT", fused)
    torch.manual_seed(7)
    model = nn.Sequential(nn.Linear(30, 70), nn.ReLU(), nn.Linear(70, 150), nn.ReLU(), nn.Linear(150, 4))
    B = 140
    g = torch.Generator().manual_seed(8)
    data = [(torch.randn(B, 30, generator=g), torch.randint(0, 4, (B,), generator=g))]
    kw = {"lr": 0.05, "momentum": 0.5, "maximize": True}
    plans, megas, _ = _ranks(model, nn.CrossEntropyLoss(), torch.optim.SGD, kw, 1, B, patience=3, data=data, grid=9)
    assert megas[0].fused_opt == (flags == "0" and fused == "1")
    emu = MegaEmulator(megas)
    emu.run(10, seed=3)
    from sparktorch_b200.early_stopper import EarlyStopping

    ref = copy.deepcopy(model)
    opt = torch.optim.SGD(ref.parameters(), **kw)
    es = EarlyStopping(patience=3)
    losses = []
    for _ in range(10):
        opt.zero_grad()
        loss = nn.CrossEntropyLoss()(ref(data[0][0]), data[0][1])
        loss.backward()
        losses.append(float(loss.detach()))
        if es.step(losses[-1]):
            break
        opt.step()
    assert len(losses) < 10
    st = emu.state[0]
    assert st["stopped"] and st["iter"] == len(losses) and st["step"] == len(losses) - 1
    np.testing.assert_allclose(st["losses"], losses, rtol=1e-4)
    _check(plans, ref)
