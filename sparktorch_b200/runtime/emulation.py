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

    def __init__(self, numel: int, spec: OptSpec, wbf16_elems: int, max_iters: int = 64, weight_dtype=torch.float32):
        self.numel, self.spec = numel, spec
        self.grad = torch.zeros(numel + TAIL, dtype=torch.float32)
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
