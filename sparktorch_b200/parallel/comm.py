"""Gradient exchange back-ends for the flat-arena trainers.

``reduce_and_update`` is the whole "distributed part" of a step: average the
gradient arena (incl. the piggy-backed loss / stop slots) over ranks and apply
the optimizer – ONE call per step instead of the reference's per-parameter
all-reduce + divide + separate ``optimizer.step()`` (distributed.py:179-198).

* :class:`LocalComm`        – world_size 1, no communication.
* :class:`CollectiveComm`   – one ``torch.distributed.all_reduce`` over the flat
  buffer (gloo on CPU; NCCL on GPU as the *oracle/baseline* path only).
* ``P2PFusedComm`` (parallel/p2p.py) – the product path on B200: our own
  sm_100a kernel doing reduce-scatter over NVLink peer loads → Adam/SGD on the
  owned shard → peer stores of the updated parameters, early-stop in-band.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from sparktorch_b200.early_stopper import EarlyStopping
from sparktorch_b200.runtime.flat import FlatArena
from sparktorch_b200.runtime.optim import OptSpec, apply_update_, new_state


class StepResult:
    __slots__ = ("loss", "stop")

    def __init__(self, loss: Optional[float], stop: bool):
        self.loss = loss
        self.stop = stop


class LocalComm:
    """Single replica.  Also the base class holding the host-side optimizer."""

    world = 1
    rank = 0
    fused = False

    def __init__(self, arena: FlatArena, spec: OptSpec, user_optimizer=None, patience: int = -1):
        self.arena = arena
        self.spec = spec
        self.user_optimizer = user_optimizer            # only for spec.kind == 'other'
        self.state: Dict[str, torch.Tensor] = new_state(spec, arena.flat_p) if spec.kind != "other" else {}
        self.step_count = 0
        self.es = EarlyStopping(patience=patience) if patience > 0 else None
        self.stopped = False

    # -- collective part (overridden) ----------------------------------------
    def _average(self) -> None:
        pass

    def broadcast_params(self) -> None:
        pass

    def average_buffers(self, module) -> None:
        pass

    # -- one step --------------------------------------------------------------
    def reduce_and_update(self, need_loss: bool) -> StepResult:
        self._average()
        loss = None
        stop = False
        if self.es is not None:
            # every rank sees the same mean loss → identical decisions; the vote
            # slot only matters for NaN divergence between ranks
            loss = float(self.arena.tail[0])
            stop = self.es.step(loss) or float(self.arena.tail[1]) > 0
        elif need_loss:
            loss = float(self.arena.tail[0])
        if stop:  # reference breaks BEFORE optimizer.step() (distributed.py:185-198)
            self.stopped = True
            return StepResult(loss, True)
        self.step_count += 1
        if self.spec.kind == "other":
            self.user_optimizer.step()
        else:
            with torch.no_grad():
                apply_update_(self.spec, self.arena.flat_p, self.arena.grads, self.state, self.step_count)
        return StepResult(loss, False)

    def finish(self) -> None:
        pass

    # -- checkpoint ------------------------------------------------------------------
    def export_state(self):
        out = {"step_count": self.step_count, "state": {k: v.detach().cpu() for k, v in self.state.items()},
               "es": self.es.state_vector() if self.es is not None else None}
        if self.user_optimizer is not None:
            out["user_optimizer"] = self.user_optimizer.state_dict()
        return out

    def import_state(self, saved) -> None:
        self.step_count = int(saved["step_count"])
        for k, v in saved["state"].items():
            self.state[k].copy_(v.to(self.state[k].device))
        if self.es is not None and saved.get("es") is not None:
            self.es.load_state_vector(saved["es"])
        if self.user_optimizer is not None and "user_optimizer" in saved:
            self.user_optimizer.load_state_dict(saved["user_optimizer"])


class CollectiveComm(LocalComm):
    def __init__(self, arena: FlatArena, spec: OptSpec, user_optimizer=None, patience: int = -1):
        super().__init__(arena, spec, user_optimizer, patience)
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def _average(self) -> None:
        dist.all_reduce(self.arena.flat_g, op=dist.ReduceOp.SUM)
        self.arena.flat_g.mul_(1.0 / self.world)

    def broadcast_params(self) -> None:
        """Rank 0's weights win (fixes the silent divergence of lazily built
        replicas)."""
        dist.broadcast(self.arena.flat_p, src=0)

    def average_buffers(self, module) -> None:
        """BatchNorm running stats etc. are averaged once at the end (quirk Q8)."""
        for b in module.buffers():
            if b.is_floating_point():
                dist.all_reduce(b, op=dist.ReduceOp.SUM)
                b.div_(self.world)
