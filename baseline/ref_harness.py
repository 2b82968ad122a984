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


class _IterationTimer:
    """Mixin: zero_grad() marks iteration boundaries (the reference calls it once per iteration)."""

    def _init_timer(self, warmup, steps, out_dir):
        self._w, self._k, self._n, self._out = int(warmup), int(steps), 0, out_dir
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
                json.dump({"device_ms": ms, "wall_ms": wall, "steps": self._k, "host_t0": self._host_t0, "host_t1": time.time()}, fh)
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
