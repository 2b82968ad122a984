"""GPU-resident hogwild: the parameter server is the NVLink-connected memory of
the GPUs (kernels in csrc/comm_hw.cu).  One barrier task per GPU; inside, workers
never wait for each other: pull (P2P all-gather of the owner shards) → forward /
backward on the local mini-batch → push (remote optimizer step on every owner's
shard).  Replaces the reference's Flask server + dill-over-HTTP transport
(hogwild.py:96-140, server.py:95-149) on a single NVSwitch domain; across nodes or
on CPU the HTTP parameter server (sparktorch_b200/server.py) is used instead.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from sparktorch_b200.ops import native
from sparktorch_b200.parallel.p2p import SymmBlock
from sparktorch_b200.runtime.engine import DeviceSampler, EagerStep, FitResult, TrainConfig, prepare_targets
from sparktorch_b200.runtime.flat import TAIL, FlatArena
from sparktorch_b200.runtime.optim import parse_optimizer
from sparktorch_b200.util import split_validation


def compensate_asynchrony(spec, world: int):
    """W asynchronous workers that apply W optimizer steps per round on gradients taken at (almost) the same stale
    parameters behave like ONE worker with extra momentum: the implicit momentum of asynchrony is 1 - 1/W (Mitliagkas et
    al., "Asynchrony begets momentum").  Stacked on Adam's beta1 = 0.9 (or SGD momentum) that is an effective momentum of
    ~0.99 at 8 workers and training at the reference's learning rates oscillates and diverges (profiles/mgc_n8.log of
    round 1: the loss falls for ~20 iterations, then climbs above ln 10).  The remote-apply push therefore runs with the
    explicit momentum reduced so that the TOTAL stays what the user asked for:  1 - beta' = min(1, W (1 - beta)).
    SPARKTORCH_HOGWILD_COMPENSATE=0 keeps the raw coefficients (the reference's behaviour)."""
    import dataclasses
    import os

    if world <= 1 or os.environ.get("SPARKTORCH_HOGWILD_COMPENSATE", "1") != "1":
        return spec
    if spec.kind in ("adam", "adamw"):
        return dataclasses.replace(spec, beta1=1.0 - min(1.0, world * (1.0 - spec.beta1)))
    if spec.kind == "sgd" and spec.momentum:
        return dataclasses.replace(spec, momentum=max(0.0, 1.0 - min(1.0, world * (1.0 - spec.momentum))))
    return spec


def pick_stripes(numel: int, world: int, max_stripes: int) -> int:
    per = (numel // 4 + world - 1) // world
    return int(max(1, min(max_stripes, per // 2048)))


def train_partition_hogwild(rank: int, world: int, model: nn.Module, criterion, opt_cls, opt_params: Optional[Dict],
                            X: np.ndarray, Y: Optional[np.ndarray], cfg: TrainConfig, device: torch.device,
                            acquire_lock: bool, window_len: int) -> FitResult:
    C = native.load(required=True)
    spec = compensate_asynchrony(parse_optimizer(opt_cls, opt_params), world)
    if spec.kind == "other":
        raise RuntimeError("GPU hogwild supports SGD/Adam/AdamW; other optimizers use the HTTP parameter server")
    t_start = time.time()
    torch.cuda.set_device(device)
    rng = np.random.default_rng(None if cfg.seed is None else cfg.seed + 7919 * rank)

    # ---- data ----------------------------------------------------------------------------
    xt = torch.from_numpy(X)
    yt = torch.from_numpy(Y) if Y is not None else None
    x_val = y_val = None
    if cfg.validation_pct > 0:
        tr, va = split_validation(X.shape[0], cfg.validation_pct, rng)
        if len(va):
            x_val = xt[torch.from_numpy(va)].to(device)
            y_val = prepare_targets(yt[torch.from_numpy(va)].to(device), criterion) if yt is not None else x_val
            xt = xt[torch.from_numpy(tr)]
            yt = yt[torch.from_numpy(tr)] if yt is not None else None
    x_all = xt.to(device)
    y_all = prepare_targets(yt.to(device), criterion) if yt is not None else x_all
    n = x_all.shape[0]

    # ---- model + arenas ------------------------------------------------------------------------
    model = model.to(device).train()
    if hasattr(criterion, "to"):
        criterion = criterion.to(device)
    # any module: Linear / Conv2d / CE / MSE on our kernels (ops/module_swap.py) - no library GEMM on the worker's step
    from sparktorch_b200.ops.module_swap import invalidate_weight_caches, swap_modules

    model, criterion, swap_report = swap_modules(model, criterion, precision=cfg.precision)
    numel = FlatArena.measure(model)
    blk = SymmBlock(device, world, rank)
    blk.reserve("param", numel * 4)
    if spec.n_state >= 1:
        blk.reserve("m", numel * 4)
    if spec.n_state >= 2:
        blk.reserve("v", numel * 4)
    blk.reserve("ctrl", C.hw_ctrl_words() * 4)
    blk.rendezvous()
    owner_param = blk.view("param", torch.float32, numel)
    local_p = torch.zeros(numel, device=device)
    local_g = torch.zeros(numel + TAIL, device=device)
    arena = FlatArena(model, device, grad_buffer=local_g, param_buffer=local_p)
    if world > 1:
        dist.broadcast(local_p, src=0)            # the job starts from rank 0's weights
    owner_param.copy_(local_p)
    loss_in = torch.zeros(2, device=device)
    mailbox = torch.zeros(4, dtype=torch.int32).pin_memory()
    stripes = pick_stripes(numel, world, C.hw_max_stripes())
    worker = C.HogwildWorker(
        world, rank, stripes, blk.peer_ptrs("param"), blk.peer_ptrs("m") if spec.n_state >= 1 else [],
        blk.peer_ptrs("v") if spec.n_state >= 2 else [], blk.peer_ptrs("ctrl"), local_p, local_g, loss_in,
        x_val is not None, numel, bool(acquire_lock), spec.kind_id, spec.lr, spec.beta1, spec.beta2, spec.eps,
        spec.weight_decay, spec.momentum, spec.dampening, spec.nesterov, spec.maximize, cfg.early_stop_patience,
        max(1, window_len), mailbox.data_ptr(), None, None, False)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()

    step = EagerStep(model, criterion, arena)
    use_mb = 0 < cfg.mini_batch < n
    sampler = DeviceSampler(n, cfg.mini_batch, device, None if cfg.seed is None else cfg.seed + rank) if use_mb else None
    loss_hist = torch.zeros(max(1, cfg.iters), device=device)
    t_loop = time.time()
    stopped_at = None
    done = 0
    for i in range(cfg.iters):
        if sampler is not None:
            idx = sampler.next()
            xb = x_all[idx]
            yb = xb if y_all is x_all else y_all[idx]
        else:
            xb, yb = x_all, y_all
        worker.pull()                              # P2P all-gather of the owners' shards
        invalidate_weight_caches()                 # (the pulled snapshot replaced the weights behind the bf16 operand copies)
        loss = step.run(xb, yb)
        loss_in[0] = loss
        if x_val is not None:
            loss_in[1] = step.evaluate(x_val, y_val)
        loss_hist[i] = loss
        worker.push()                              # remote optimizer step on every owner's shard
        done = i + 1
        if cfg.verbose:
            print(f"Partition: {rank}. Iteration: {i}. Loss: {float(loss)}, "
                  f"Val Loss: {float(loss_in[1]) if x_val is not None else None}")
        if cfg.early_stop_patience > 0:
            torch.cuda.current_stream(device).synchronize()
            if int(mailbox[0]) != 0:
                stopped_at = i
                break
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()                             # every worker has pushed its last gradient
    worker.pull()
    torch.cuda.synchronize(device)
    t_end = time.time()
    ctrl0 = blk.view("ctrl", torch.int32, C.hw_ctrl_words())
    history = {
        "engine": "eager", "kernel_swap": swap_report, "comm": "HogwildP2P" + ("+locks" if acquire_lock else ""), "world_size": world,
        "train_loss": loss_hist[:done].cpu().tolist(), "stopped_at": stopped_at, "iters_run": done,
        "samples_per_step": (cfg.mini_batch if use_mb else n), "stripes": stripes,
        "pushes_total": int(ctrl0[2 * C.hw_max_stripes() + 1]) if rank == 0 else None,
        "loop_seconds": t_end - t_loop, "total_seconds": t_end - t_start,
    }
    state = arena.detach_state_dict() if rank == 0 else None
    if world > 1:
        dist.barrier()                             # keep peer memory alive until everyone has read it
    return FitResult(state, history)


# ----------------------------------------------------------------------------- fused (graph-captured) worker
class HogwildBackend:
    """Stands in for ``FusedUpdate`` inside :class:`FusedMLP`: parameters live sharded on their owners
    (symmetric memory), gradients / bf16 weight copies / step counters are private to the worker."""

    use_multicast = False

    def __init__(self, numel, wbf16_elems, spec, world, rank, device, patience, window_len, max_iters, acquire_lock):
        C = native.load(required=True)
        self.C, self.numel, self.world, self.rank, self.device = C, numel, world, rank, device
        blk = SymmBlock(device, world, rank)
        blk.reserve("param", numel * 4)
        if spec.n_state >= 1:
            blk.reserve("m", numel * 4)
        if spec.n_state >= 2:
            blk.reserve("v", numel * 4)
        blk.reserve("ctrl", C.hw_ctrl_words() * 4)
        blk.rendezvous()
        self.block = blk
        self.owner_param = blk.view("param", torch.float32, numel)
        self.param = torch.zeros(numel, device=device)                  # local snapshot (init + final pull)
        self.grad = torch.zeros(numel + TAIL, device=device)
        self.wbf16 = torch.zeros(wbf16_elems, dtype=torch.bfloat16, device=device)
        self.state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=device)
        self.loss_hist = torch.zeros(max(1, max_iters), device=device)
        self.mailbox = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.loss_in = torch.zeros(2, device=device)
        n4 = numel // 4
        self.shard_elems = ((n4 + world - 1) // world) * 4
        self.peer_param_ptrs = blk.peer_ptrs("param")
        self.peer_table = torch.tensor(self.peer_param_ptrs, dtype=torch.int64, device=device)   # read by the K-AG GEMMs
        self.stripes = pick_stripes(numel, world, C.hw_max_stripes())
        self.updater = C.HogwildWorker(
            world, rank, self.stripes, blk.peer_ptrs("param"), blk.peer_ptrs("m") if spec.n_state >= 1 else [],
            blk.peer_ptrs("v") if spec.n_state >= 2 else [], blk.peer_ptrs("ctrl"), self.param, self.grad, self.loss_in,
            False, numel, bool(acquire_lock), spec.kind_id, spec.lr, spec.beta1, spec.beta2, spec.eps, spec.weight_decay,
            spec.momentum, spec.dampening, spec.nesterov, spec.maximize, patience, max(1, window_len),
            self.mailbox.data_ptr(), self.state, self.loss_hist, True)

    def iters_done(self) -> int:
        return int(self.mailbox[1])

    def stopped(self) -> bool:
        return bool(self.mailbox[0])

    def last_loss(self) -> float:
        return float(self.mailbox[2:3].view(torch.float32)[0])


def train_partition_hogwild_fused(rank, world, model, criterion, opt_cls, opt_params, X, Y, cfg: TrainConfig, device,
                                  window_len: int, acquire_lock: bool = False):
    """Lock-free hogwild for lowered MLPs: the whole worker iteration is one CUDA graph -
    K-AG forward GEMMs (weights gathered from their owners inside the GEMM), fused loss, backward GEMMs,
    remote-apply push - replayed by the native runner without Python or barriers.  ``acquire_lock`` (the reference's default,
    torch_distributed.py:189): the push applies every stripe under its owner's spin lock, so concurrent updates of a stripe
    are serialised like the server's POST /update; the K-AG gather stays a wait-free read (it may see elements of two
    consecutive versions of a stripe - each element itself is never torn).  Returns None when the model does not lower
    (caller falls back to the eager hogwild worker, whose GEMMs run on our kernels through ops/module_swap.py)."""
    from sparktorch_b200.runtime.fused_mlp import FusedMLP, _pad, _to_bf16_padded, loss_kind, lower_mlp

    spec = compensate_asynchrony(parse_optimizer(opt_cls, opt_params), world)
    if spec.kind == "other" or cfg.precision not in ("auto", "bf16"):
        return None
    lowered = lower_mlp(model)
    is_cnn = False
    if lowered is None:
        from sparktorch_b200.runtime.fused_cnn import FusedCNN, lower_cnn

        lowered = lower_cnn(model, X.shape[1])
        is_cnn = lowered is not None
    kind = loss_kind(criterion, lowered) if lowered is not None else None
    if kind is None or (Y is None and kind != "mse") or cfg.validation_pct > 0:
        return None
    last_out = (lowered.linears if is_cnn else lowered.layers)[-1].module.out_features
    if last_out > 32:
        return None
    C = native.load(required=True)
    t_start = time.time()
    torch.cuda.set_device(device)
    if Y is None:
        Y = X
    n = X.shape[0]
    use_mb = 0 < cfg.mini_batch < n
    Bt = cfg.mini_batch if use_mb else n
    d0 = X.shape[1]
    model = model.to(device)
    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        factory = lambda numel, wb: HogwildBackend(numel, wb, spec, world, rank, device, cfg.early_stop_patience,  # noqa: E731
                                                   window_len, cfg.iters + 1, bool(acquire_lock))
        if is_cnn:
            plan = FusedCNN(model, lowered, kind, spec, world, rank, device, Bt, 0, cfg.early_stop_patience, cfg.iters + 1,
                            seed=0 if cfg.seed is None else int(cfg.seed), backend_factory=factory)
        else:
            plan = FusedMLP(model, lowered, kind, spec, world, rank, device, Bt, 0, cfg.early_stop_patience, cfg.iters + 1,
                            backend_factory=factory)
        be = plan.fu
        if world > 1:
            dist.broadcast(be.param, src=0)
        be.owner_param.copy_(be.param)
        src = _to_bf16_padded(torch.from_numpy(np.ascontiguousarray(X)), device)
        yt = torch.from_numpy(np.ascontiguousarray(Y))
        tgt = (yt.reshape(-1).to(torch.int32) if kind == "ce" else yt.to(torch.float32).reshape(n, -1)).to(device)
        plan.train["acts"][0], plan.train["tgt"] = src, tgt
        windowed = use_mb
        if windowed:
            plan.window = (max(1, n // Bt), Bt)
        plan.loss_slot = be.loss_in[0:1]
        plan.train_loss_slot = be.loss_in[0:1]

        def enqueue():
            plan.n_launch = 0
            plan.forward(plan.train, slot=plan.loss_slot, train=True, windowed=windowed)
            plan.backward(plan.train, windowed=windowed)
            be.updater.step()
            plan.n_launch += 1

        # warm-up outside capture without touching the shared parameters: forward/backward + a dry push
        plan.forward(plan.train, slot=plan.loss_slot, train=True, windowed=windowed)
        plan.backward(plan.train, windowed=windowed)
        be.updater.warm()
        stream.synchronize()
        be.grad.zero_()
        be.loss_in.zero_()
        if world > 1:
            dist.barrier()
        runner = C.GraphRunner()
        use_pdl = True
        for attempt in range(2):
            C.set_pdl(use_pdl)
            try:
                runner.begin_capture()
                enqueue()
                runner.end_capture()
                break
            except RuntimeError:
                C.set_pdl(False)
                if attempt == 1:
                    raise
                try:
                    runner.end_capture()
                except RuntimeError:
                    pass
                runner, use_pdl = C.GraphRunner(), False
        C.set_pdl(False)
        from sparktorch_b200.runtime.engine import WindowTimer

        window = WindowTimer(cfg.time_window, device, world)
        check = 4 if cfg.early_stop_patience > 0 else 0
        t_loop = time.time()
        segs = [(0, cfg.iters)]
        if cfg.time_window is not None:
            w0, wk = cfg.time_window
            segs = [(0, w0), (w0, w0 + wk), (w0 + wk, cfg.iters)]
        for si, (lo, hi) in enumerate(segs):
            if hi <= lo:
                continue
            timed = cfg.time_window is not None and si == 1
            if timed:
                window.before_step(lo)
            runner.run(lo, hi - lo, be.mailbox.data_ptr(), check)
            if timed:
                window.after_step(hi - 1)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        be.updater.pull()
        stream.synchronize()
    t_end = time.time()
    done = be.iters_done()
    history = {
        "engine": "fused_cnn" if is_cnn else "fused_mlp", "comm": "HogwildP2P+K-AG" + ("+locks" if acquire_lock else ""),
        "world_size": world, "precision": "bf16",
        "window_ms": window.result_ms(),
        "train_loss": be.loss_hist[:done].cpu().tolist(), "stopped_at": (done - 1) if be.stopped() else None,
        "iters_run": done, "samples_per_step": Bt, "stripes": be.stripes, "gpu_launches_per_step": plan.n_launch,
        "pdl": use_pdl, "loop_seconds": t_end - t_loop, "total_seconds": t_end - t_start,
        "pushes_total": int(be.block.view("ctrl", torch.int32, C.hw_ctrl_words())[2 * C.hw_max_stripes() + 1]) if rank == 0 else None,
    }
    state = plan.arena.detach_state_dict() if rank == 0 else None
    if world > 1:
        dist.barrier()
    del _pad
    return FitResult(state, history)
