"""CPU emulation of the device-side pieces the fused plans depend on.

``EmulatedUpdate`` stands in for :class:`parallel.p2p.FusedUpdate` (symmetric arenas + the K-AR kernel): same
attributes (``grad`` / ``param`` / ``wbf16`` / ``state`` / ``updater``), same segment table, but a single replica in
host memory whose ``step`` is the plain optimizer of ``runtime/optim.py``.  Together with ``ops/reference.py`` it lets
the *unchanged* plan code (``FusedResNet`` ...) run in fp32 on the CPU, where forward, backward and update can be
compared exactly with autograd + ``torch.optim`` (tests/test_plan_emulation.py).
"""
from __future__ import annotations

from typing import List

import torch

from sparktorch_b200.runtime.flat import TAIL
from sparktorch_b200.runtime.optim import OptSpec, apply_update_, new_state


class _EmulatedUpdater:
    def __init__(self, owner: "EmulatedUpdate"):
        self.o = owner

    def publish(self) -> None:
        o = self.o
        for p_off, rows, cols, w_off, ldw, _wt_off, _wt_ld, _pub in o.segments:
            if w_off >= 0:
                src = o.param[p_off: p_off + rows * cols].view(rows, cols)
                o.wbf16[w_off: w_off + rows * ldw].view(rows, ldw)[:, :cols] = src.to(o.wbf16.dtype)

    def step(self) -> None:
        o = self.o
        o.steps += 1
        apply_update_(o.spec, o.param, o.grad[: o.numel], o.opt_state, o.steps)
        o.loss_hist[o.steps - 1] = o.grad[o.numel]
        o.mailbox[1] = o.steps
        o.state[4] = o.steps                  # StepState.iter: the batch window of the next step
        o.grad.zero_()
        self.publish()


class EmulatedUpdate:
    wants_segments = True          # plans built through ``backend_factory`` must still hand over their segment table
    use_multicast = False
    exchange = "emulated"

    def __init__(self, numel: int, spec: OptSpec, wbf16_elems: int, max_iters: int = 64, weight_dtype=torch.float32,
                 grad_extra: int = 0):
        self.numel, self.spec = numel, spec
        self.grad = torch.zeros(numel + TAIL + (60 + grad_extra if grad_extra else 0), dtype=torch.float32)
        self.param = torch.zeros(numel, dtype=torch.float32)
        self.wbf16 = torch.zeros(max(1, wbf16_elems), dtype=weight_dtype)
        self.state = torch.zeros(16, dtype=torch.int32)
        self.loss_hist = torch.zeros(max_iters, dtype=torch.float32)
        self.mailbox = torch.zeros(4, dtype=torch.int32)
        self.opt_state = new_state(spec, self.param)
        self.m = self.opt_state.get("m", self.opt_state.get("buf"))
        self.v = self.opt_state.get("v")
        self.segments: List[List[int]] = []
        self.steps = 0
        self.updater = _EmulatedUpdater(self)

    def build(self, segment_rows: List[List[int]], zero_grads: bool = True) -> None:
        self.segments = [list(r) for r in segment_rows]

    def iters_done(self) -> int:
        return self.steps

    def stopped(self) -> bool:
        return False

    def last_loss(self) -> float:
        return float(self.loss_hist[max(0, self.steps - 1)])


class EmulatedHogwild(EmulatedUpdate):
    """Single-process stand-in for ``parallel.hogwild_p2p.HogwildBackend``: W notional owners, each holding an arena in
    which only its own shard is authoritative (the rest is poisoned with NaN, so a gather from the wrong owner is
    caught); the forward GEMMs gather weights through ``peer_table`` (K-AG), ``step`` applies the update to the owners'
    shards (the remote-apply push with one worker)."""

    def __init__(self, numel: int, spec: OptSpec, wbf16_elems: int, world: int = 3, max_iters: int = 64):
        super().__init__(numel, spec, wbf16_elems, max_iters)
        self.world = world
        n4 = numel // 4
        self.shard_elems = ((n4 + world - 1) // world) * 4
        self.peer_table = [torch.full((numel,), float("nan")) for _ in range(world)]
        self.updater = _EmulatedHogwildWorker(self)

    def scatter(self) -> None:
        """owners <- the local snapshot (initial weights)"""
        for r, t in enumerate(self.peer_table):
            lo, hi = r * self.shard_elems, min(self.numel, (r + 1) * self.shard_elems)
            if lo < hi:
                t[lo:hi] = self.param[lo:hi]

    def gather(self) -> None:
        """local snapshot <- owners (the final pull)"""
        for r, t in enumerate(self.peer_table):
            lo, hi = r * self.shard_elems, min(self.numel, (r + 1) * self.shard_elems)
            if lo < hi:
                self.param[lo:hi] = t[lo:hi]


class _EmulatedHogwildWorker(_EmulatedUpdater):
    def publish(self) -> None:          # weights are gathered inside the forward GEMMs, nothing to build up front
        self.o.scatter()

    def step(self) -> None:
        o = self.o
        o.gather()
        o.steps += 1
        apply_update_(o.spec, o.param, o.grad[: o.numel], o.opt_state, o.steps)
        o.loss_hist[o.steps - 1] = o.grad[o.numel]
        o.mailbox[1] = o.steps
        o.state[4] = o.steps
        o.grad.zero_()
        o.scatter()
