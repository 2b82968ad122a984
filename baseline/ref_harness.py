"""Harness pieces for running the UNMODIFIED reference package
(``baseline/_ref``, installed with pip from /root/reference) on this box.

Nothing here is on the product path.  It provides:

* :func:`install_backend_patch` – the reference hard-codes
  ``init_process_group('gloo')`` (distributed.py:110).  BASELINE.md defines the
  bar as the reference's own loop with gloo→nccl, so for CUDA runs the backend
  argument is swapped at the call boundary (env ``SPARKTORCH_REF_BACKEND=gloo``
  keeps the stock transport).  The reference's files are not touched.
* :class:`TimedAdam` – an ``Adam`` subclass the *user-level API* lets us pass in
  the TorchObj; its ``zero_grad`` (called once at the top of every reference
  iteration, distributed.py:142) records CUDA events so exactly K iterations
  can be device-timed without editing the reference loop.
"""
from __future__ import annotations

import json
import os
import time

import torch


def install_backend_patch() -> None:
    import torch.distributed as dist

    want = os.environ.get("SPARKTORCH_REF_BACKEND", "nccl")
    if want == "gloo" or getattr(dist.init_process_group, "_ref_patched", False):
        return
    orig = dist.init_process_group

    def patched(backend=None, *args, **kwargs):
        if backend == "gloo" and torch.cuda.is_available():
            backend = "nccl"
            kwargs.setdefault("device_id", torch.device("cuda", torch.cuda.current_device()))
        return orig(backend, *args, **kwargs)

    patched._ref_patched = True  # type: ignore[attr-defined]
    dist.init_process_group = patched


LOSS_LOG: list = []       # per process: every loss value the reference loop computed (see LoggedCrossEntropy)


class LoggedCrossEntropy(torch.nn.CrossEntropyLoss):
    """User-level criterion (travels in the TorchObj like any other): identical math, remembers the loss values so the
    bench can show that the reference arm is learning too.  The reference's first call per iteration raises for float
    labels (distributed.py:152-157) before anything is logged; the retry with long labels is the one recorded."""

    def forward(self, input, target):      # noqa: A002
        out = super().forward(input, target)
        if out.requires_grad:              # the training loss (validation runs detached in the reference as well, but is rare here)
            LOSS_LOG.append(out.detach())
        return out


def loss_summary():
    if not LOSS_LOG:
        return None
    vals = torch.stack([v.float().reshape(()) for v in LOSS_LOG]).cpu().tolist()
    k = max(1, min(3, len(vals) // 2))
    return {"first": sum(vals[:k]) / k, "last": sum(vals[-k:]) / k, "n": len(vals)}


class _IterationTimer:
    """Mixin: zero_grad() marks iteration boundaries (the reference calls it once per iteration)."""

    def _init_timer(self, warmup, steps, out_dir):
        self._w, self._k, self._n, self._out = int(warmup), int(steps), 0, out_dir
        LOSS_LOG.clear()
        self._e0 = self._e1 = None
        self._t0 = None

    def _fence(self):
        import torch.distributed as dist

        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
            if torch.cuda.is_available():
                torch.cuda.synchronize()

    def zero_grad(self, set_to_none: bool = True):
        i = self._n
        self._n += 1
        cuda = torch.cuda.is_available()
        if i == self._w:
            self._fence()
            if cuda:
                self._e0 = torch.cuda.Event(enable_timing=True)
                self._e1 = torch.cuda.Event(enable_timing=True)
            self._t0 = time.perf_counter()
            self._host_t0 = time.time()
            if cuda:
                self._e0.record()
        elif i == self._w + self._k:
            if cuda:
                self._e1.record()
                torch.cuda.synchronize()
            wall = (time.perf_counter() - self._t0) * 1e3
            ms = self._e0.elapsed_time(self._e1) if cuda else wall
            import torch.distributed as dist

            rank = dist.get_rank() if dist.is_initialized() else 0
            with open(os.path.join(self._out, f"ref_timing_rank{rank}.json"), "w") as fh:
                json.dump({"device_ms": ms, "wall_ms": wall, "steps": self._k, "host_t0": self._host_t0, "host_t1": time.time(),
                           "loss": loss_summary()}, fh)
            self._fence()
        return super().zero_grad(set_to_none)


class TimedAdam(_IterationTimer, torch.optim.Adam):
    def __init__(self, params, warmup=3, steps=10, out_dir="/tmp", **kw):
        torch.optim.Adam.__init__(self, params, **kw)
        self._init_timer(warmup, steps, out_dir)


class TimedSGD(_IterationTimer, torch.optim.SGD):
    def __init__(self, params, warmup=3, steps=10, out_dir="/tmp", **kw):
        torch.optim.SGD.__init__(self, params, **kw)
        self._init_timer(warmup, steps, out_dir)


# ------------------------------------------------------------------------------------------------ hogwild baseline
def hogwild_nccl_worker(rank: int, world: int, device, net, X, Y, iters: int, warmup: int, steps: int, lr: float, batch: int):
    """The reference's hogwild data flow (hogwild.py:96-140 worker, server.py:95-149 parameter server) with its transport -
    dill over HTTP to a CPU Flask process, which ignores `device` - replaced by NCCL on the GPUs, as BASELINE.md B2 defines the
    bar: rank 0 plays the server (owns the parameters and the only optimizer); every iteration each worker
      pull : dist.broadcast(param, src=0) for every parameter          (GET /parameters + load_state_dict, hogwild.py:103-104)
      step : the reference's forward / criterion / backward on its mini-batch, torch eager fp32 (cuBLAS / cuDNN)
      push : dist.reduce(grad, dst=0) for every parameter, server-side optimizer.step()   (POST /update, server.py:127-149)
    Returns (device_ms over `steps` iterations, wall_ms, first/last loss)."""
    import torch.distributed as dist
    import torch.nn as nn

    net = net.to(device).train()
    crit = nn.CrossEntropyLoss()
    opt = torch.optim.Adam(net.parameters(), lr=lr) if rank == 0 else None
    x_all = torch.from_numpy(X).to(device)
    y_all = torch.from_numpy(Y).to(device).reshape(-1).long()
    n = x_all.shape[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    t0 = host0 = host1 = wall = None

    def fence():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    for i in range(iters):
        if i == warmup:
            fence()
            t0, host0 = time.perf_counter(), time.time()
            e0.record()
        elif i == warmup + steps:
            e1.record()
            torch.cuda.synchronize(device)
            wall, host1 = (time.perf_counter() - t0) * 1e3, time.time()
            fence()
        if world > 1:
            for p_ in net.parameters():
                dist.broadcast(p_.data, src=0)
        idx = torch.randperm(n, device=device, generator=gen)[:batch] if batch < n else None
        xb, yb = (x_all, y_all) if idx is None else (x_all[idx], y_all[idx])
        net.zero_grad(set_to_none=False)          # (the reference forgets this - quirk Q9 - which only changes the numbers)
        loss = crit(net(xb), yb)
        losses.append(float(loss.item()))         # the reference calls loss.item() every iteration (hogwild.py:113)
        loss.backward()
        if world > 1:
            for p_ in net.parameters():
                dist.reduce(p_.grad, dst=0)
        if rank == 0:
            opt.step()
    if wall is None:                               # iters == warmup + steps: close the window here
        e1.record()
        torch.cuda.synchronize(device)
        wall, host1 = (time.perf_counter() - t0) * 1e3, time.time()
    k = max(1, min(3, len(losses) // 2))
    return {"device_ms": e0.elapsed_time(e1), "wall_ms": wall, "host_t0": host0, "host_t1": host1,
            "loss": {"first": sum(losses[:k]) / k, "last": sum(losses[-k:]) / k, "n": len(losses)}}
