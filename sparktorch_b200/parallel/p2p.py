"""Symmetric (peer-mapped) device memory + the fused all-reduce/optimizer comm.

Every rank allocates the same block layout; after :meth:`SymmBlock.rendezvous`
each rank holds the device addresses of all peers' blocks (CUDA VMM handles
exchanged through ``torch.distributed._symmetric_memory``) and, when the fabric
supports it, one NVLS multicast address aliasing all of them.  Our kernels then
``ld``/``st`` peer memory directly over NVLink 5 – no NCCL on the training path
(NCCL only bootstraps the handle exchange and serves as test oracle).

:class:`P2PFusedComm` plugs the K-AR kernel (csrc/comm_ar.cu) into the generic
flat-arena engine: one launch per step does reduce-scatter → Adam/SGD → publish,
with the loss / early-stop decision in-band.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from sparktorch_b200.ops import native
from sparktorch_b200.runtime.flat import TAIL, FlatArena
from sparktorch_b200.runtime.optim import OptSpec

_ALIGN = 1024


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class SymmBlock:
    """One symmetric allocation carved into named regions."""

    def __init__(self, device: torch.device, world: int, rank: int):
        self.device, self.world, self.rank = device, world, rank
        self._plan: List[Tuple[str, int]] = []
        self._off: Dict[str, int] = {}
        self.total = 0
        self.buf: Optional[torch.Tensor] = None
        self.peer_base: List[int] = []
        self.multicast_base = 0
        self._handle = None

    def reserve(self, name: str, nbytes: int) -> None:
        self._off[name] = self.total
        self._plan.append((name, nbytes))
        self.total += _round_up(max(nbytes, 16), _ALIGN)

    def rendezvous(self) -> None:
        if self.world == 1:
            self.buf = torch.zeros(self.total, dtype=torch.uint8, device=self.device)
            self.peer_base = [self.buf.data_ptr()]
            return
        import torch.distributed._symmetric_memory as symm

        group = dist.group.WORLD
        self.buf = symm.empty(self.total, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self._handle = symm.rendezvous(self.buf, group.group_name)
        self.peer_base = [int(p) for p in self._handle.buffer_ptrs]
        try:
            self.multicast_base = int(getattr(self._handle, "multicast_ptr", 0) or 0)
        except Exception:
            self.multicast_base = 0
        torch.cuda.synchronize(self.device)
        dist.barrier()

    def view(self, name: str, dtype: torch.dtype, numel: int) -> torch.Tensor:
        off = self._off[name]
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self.buf[off: off + nbytes].view(dtype)

    def peer_ptrs(self, name: str) -> List[int]:
        return [b + self._off[name] for b in self.peer_base]

    def multicast_ptr(self, name: str) -> int:
        return self.multicast_base + self._off[name] if self.multicast_base else 0


def opt_state_elems(numel: int, world: int) -> int:
    """Optimizer state is kept for the owned shard only (ZeRO-1)."""
    n4 = numel // 4
    per = (n4 + world - 1) // world
    return per * 4


class FusedUpdate:
    """Owns the symmetric arenas + the native ``Updater`` for one model replica."""

    def __init__(self, numel: int, spec: OptSpec, world: int, rank: int, device: torch.device, patience: int,
                 max_iters: int, wbf16_elems: int = 0, use_multicast: bool = True, exchange: str = "twoshot",
                 grad_extra: int = 0):
        C = native.load(required=True)
        self.C = C
        self.numel, self.world, self.rank, self.device = numel, world, rank, device
        # 'push': one-shot exchange - the GEMM epilogues mirror gradient writes into receive slots on every peer
        # (double buffered by iteration parity); optimizer state is replicated.  'twoshot': reduce-scatter/publish.
        self.exchange = exchange if world > 1 else "twoshot"
        self.recv_stride = _round_up(numel + TAIL, 64)
        self.block = SymmBlock(device, world, rank)
        self.grad_extra = grad_extra      # K-STEP partial slots (bias-gradient / loss partials) behind the tail
        self.block.reserve("grad", (numel + TAIL + 60 + grad_extra) * 4)
        if self.exchange == "push":
            self.block.reserve("recv", 2 * world * self.recv_stride * 4)
        self.block.reserve("param", numel * 4)
        if wbf16_elems:
            self.block.reserve("wbf16", wbf16_elems * 2)
        self.block.reserve("flags", C.flag_words() * 4)
        self.block.rendezvous()
        self.grad = self.block.view("grad", torch.float32, numel + TAIL + (60 + grad_extra if grad_extra else 0))
        self.param = self.block.view("param", torch.float32, numel)
        self.wbf16 = self.block.view("wbf16", torch.bfloat16, wbf16_elems) if wbf16_elems else None
        n_state = numel if self.exchange == "push" else opt_state_elems(numel, world)
        self.m = torch.zeros(n_state, device=device) if spec.n_state >= 1 else None
        self.v = torch.zeros(n_state, device=device) if spec.n_state >= 2 else None
        self.state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=device)
        self.loss_hist = torch.zeros(max(1, max_iters), dtype=torch.float32, device=device)
        self.mailbox = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.spec, self.patience = spec, patience
        import os

        self.use_multicast = (use_multicast and world > 1 and self.block.multicast_base != 0
                              and os.environ.get("SPARKTORCH_NVLS", "1") == "1")
        self.updater = None
        self._segs = None

    def build(self, segment_rows: List[List[int]], zero_grads: bool = True) -> None:
        C, s = self.C, self.spec
        self._segs = torch.frombuffer(bytearray(C.pack_segments(segment_rows)), dtype=torch.uint8).to(self.device)
        mc = self.block.multicast_ptr("grad") if self.use_multicast else 0
        flags = self.block.peer_ptrs("flags") if self.world > 1 else []
        self.updater = C.Updater(
            self.world, self.rank, self.block.peer_ptrs("grad"), self.block.peer_ptrs("param"),
            self.block.peer_ptrs("wbf16") if self.wbf16 is not None else [], flags, mc, self.m, self.v, self.state,
            self.loss_hist, self.mailbox.data_ptr(), self._segs, self.numel, TAIL, s.kind_id, s.lr, s.beta1, s.beta2,
            s.eps, s.weight_decay, s.momentum, s.dampening, s.nesterov, s.maximize, self.patience, zero_grads,
            self.block.peer_ptrs("recv")[self.rank] if self.exchange == "push" else 0, self.recv_stride,
            # NVLS publication of the updated values (SPARKTORCH_NVLS_PUBLISH=0: per-peer stores as in round 1)
            self.block.multicast_ptr("param") if (self.use_multicast and self._mc_publish()) else 0,
            self.block.multicast_ptr("wbf16") if (self.use_multicast and self._mc_publish() and self.wbf16 is not None) else 0)

    @staticmethod
    def _mc_publish() -> bool:
        import os

        return os.environ.get("SPARKTORCH_NVLS_PUBLISH", "1") == "1"

    def push_kw(self, mask: int) -> dict:
        """Arguments that make a GEMM mirror its gradient-arena writes into my slot on every peer."""
        if self.exchange != "push" or mask == 0:
            return {}
        W, me, st = self.world, self.rank, self.recv_stride
        local = self.block.peer_ptrs("grad")[me]
        peers = self.block.peer_ptrs("recv")
        if getattr(self, "_push_table", None) is None:
            deltas = [peers[p] + ((par * W + me) * st) * 4 - local for par in range(2) for p in range(W)]
            self._push_table = torch.tensor(deltas, dtype=torch.int64, device=self.device)
        return dict(push_table=self._push_table, push_me=me, push_mask=mask, push_state=self.state)

    # StepState accessors (int32 view: epoch, blocks_done, step(lo,hi), iter(lo,hi), stopped, ...)
    def iters_done(self) -> int:
        return int(self.mailbox[1])

    def stopped(self) -> bool:
        return bool(self.mailbox[0])

    def last_loss(self) -> float:
        return float(self.mailbox[2:3].view(torch.float32)[0])


class P2PFusedComm:
    """Comm back-end of the eager engine on GPUs (any nn.Module, fp32 params):
    gradients of autograd land in the symmetric arena, one K-AR launch per step."""

    fused = True

    def __init__(self, model: torch.nn.Module, spec: OptSpec, patience: int, world: int, rank: int,
                 device: torch.device, max_iters: int):
        numel = FlatArena.measure(model)
        self.fu = FusedUpdate(numel, spec, world, rank, device, patience, max_iters)
        self.arena = FlatArena(model, device, grad_buffer=self.fu.grad, param_buffer=self.fu.param)
        rows = [[o, 1, p.numel(), -1, 0, -1, 0, 1] for p, o in zip(self.arena.params, self.arena.offsets)]
        # autograd overwrites .grad views in place, the kernel clears the arena after each exchange
        self.fu.build(rows, zero_grads=True)
        self.world, self.rank = world, rank
        self.spec = spec
        self.patience = patience
        self.stopped = False
        self.step_count = 0

    def broadcast_params(self) -> None:
        if self.world > 1:
            dist.broadcast(self.arena.flat_p, src=0)

    def average_buffers(self, module) -> None:
        if self.world > 1:
            for b in module.buffers():
                if b.is_floating_point():
                    dist.all_reduce(b, op=dist.ReduceOp.SUM)
                    b.div_(self.world)

    def reduce_and_update(self, need_loss: bool):
        from sparktorch_b200.parallel.comm import StepResult

        self.fu.updater.step()
        self.step_count += 1
        if need_loss or self.patience > 0:
            torch.cuda.current_stream(self.fu.device).synchronize()
            stop = self.fu.stopped()
            self.stopped = stop
            return StepResult(self.fu.last_loss(), stop)
        return StepResult(None, False)

    def after_graph_step(self, need_loss: bool):
        """Bookkeeping of a step whose K-AR launch was part of a replayed CUDA graph (runtime/engine.py GraphedStep)."""
        from sparktorch_b200.parallel.comm import StepResult

        self.step_count += 1
        if need_loss or self.patience > 0:
            torch.cuda.current_stream(self.fu.device).synchronize()
            stop = self.fu.stopped()
            self.stopped = stop
            return StepResult(self.fu.last_loss(), stop)
        return StepResult(None, False)

    def finish(self) -> None:
        torch.cuda.synchronize(self.fu.device)

    # -- checkpoint: optimizer shard + device-side step state ------------------------------------
    def export_state(self):
        fu = self.fu
        return {"m": None if fu.m is None else fu.m.cpu(), "v": None if fu.v is None else fu.v.cpu(),
                "state": fu.state.cpu(), "loss_hist": fu.loss_hist.cpu(), "step_count": self.step_count}

    def import_state(self, saved) -> None:
        fu = self.fu
        if fu.m is not None and saved["m"] is not None:
            fu.m.copy_(saved["m"].to(fu.device))
        if fu.v is not None and saved["v"] is not None:
            fu.v.copy_(saved["v"].to(fu.device))
        st = saved["state"].to(fu.device)
        fu.state[2:].copy_(st[2:])          # keep this run's barrier epoch, restore step/iter/early-stop record
        fu.loss_hist.copy_(saved["loss_hist"].to(fu.device))
        self.step_count = int(saved["step_count"])
