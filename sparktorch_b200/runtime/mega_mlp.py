"""K-STEP: whole training steps of a Linear/ReLU chain as ONE persistent kernel launch.

``csrc/mlp_step.cu`` executes a static task graph; this module *builds* that graph for a lowered MLP
(``runtime/fused_mlp.py``) and places it on the SMs:

    F_l    forward tiles      A_l = relu(A_{l-1} W_l^T + b_l)         (head: + fused softmax-CE / MSE -> dZ_L, db_L, loss)
    D      decision           mean loss over the replicas, early stopping, loss history, host mailbox
    B_l    dgrad tiles        dZ_{l-1} = (dZ_l W_l) * (A_{l-1} > 0), column sums -> db_{l-1}
    G_l    wgrad tiles        dW_l = dZ_l^T A_{l-1}  -> symmetric gradient arena (plain stores)
    U_s    update chunks      owner reduces its shard over the replicas (NVLS / peer loads), Adam/SGD,
                              publishes fp32 + bf16 to every replica
    H_L    quiet head copies  (chained head) the B_L tiles' CTAs recompute logits / dZ_L of their rows and hand dZ_L to the
                              dgrad MMA through shared memory; nothing is stored or signalled - F_L keeps that job

Dependencies are monotonic counters (inside a GPU) and epoch flags (between GPUs); see the kernel's header
comment.  The reference loop this replaces: /root/reference/sparktorch/distributed.py:140-203.

Everything here is plain Python/NumPy over tensor addresses, so the *same* builder runs on CPU tensors:
``MegaEmulator`` interprets the task tables of W emulated ranks under a randomised scheduler and is compared
against autograd + torch.optim in tests/test_mega_plan.py - table bugs (tile coverage, windows, partial
slots, ownership, flag masks, dead-locks, cross-step hazards) are found without a GPU.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from sparktorch_b200.runtime.flat import TAIL

MS_MAX_TASKS = 24
MS_MAX_FLAGS = 64
MS_GEMM, MS_UPDATE, MS_DECIDE = 1, 2, 3
F_A_MN, F_B_MN, F_RELU, F_A_WIN, F_B_WIN, F_BIAS, F_MASK = 1, 2, 4, 8, 16, 32, 64
F_OUT_BF16, F_OUT_F32, F_COLSUM, F_LOSS_CE, F_LOSS_MSE, F_LBL_WIN, F_FUSED_OPT = 128, 256, 512, 1024, 2048, 4096, 8192
F_CHAIN_HEAD, F_QUIET, F_CHAIN_TAIL, F_EARLY_A = 1 << 14, 1 << 15, 1 << 16, 1 << 17
CNT_STRIDE = 32            # words between dependency counters on the device (MS_CNT_STRIDE in csrc/mlp_step.h)
XCNT_BASE = 1024           # word offset of the cross-GPU counters in the symmetric flag block (MS_XCNT_BASE)

# mirrors struct MsTask (csrc/mlp_step.h), 256 bytes
MS_TASK = np.dtype([
    ("type", "<i4"), ("flags", "<i4"), ("tmap_a", "<i4", 2), ("tmap_b", "<i4", 2),
    ("m0", "<i4"), ("n0", "<i4"), ("M", "<i4"), ("N", "<i4"), ("num_kb", "<i4"), ("block_n", "<i4"),
    ("ld_mask", "<i4"), ("ld_out", "<i4"), ("ld_f32", "<i4"), ("ld_target", "<i4"),
    ("bias", "<i8"), ("mask", "<i8", 2), ("out_bf16", "<i8", 2), ("out_f32", "<i8"), ("colsum", "<i8"),
    ("labels", "<i8"), ("loss_slot", "<i8", 2), ("loss_scale", "<f4"), ("seg", "<i4"), ("e0", "<i8"),
    ("len", "<i4"), ("n_part", "<i4"), ("part_off", "<i8"), ("part_stride", "<i8"),
    ("wait", "<i4", (3, 4)), ("sig", "<i4", 4), ("ep_c0", "<i4"), ("ep_cols", "<i4")])
assert MS_TASK.itemsize == 256


def _pad(n: int, m: int = 8) -> int:
    return (n + m - 1) // m * m


def _cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


def mega_extra_elems(d: Sequence[int], has_bias: Sequence[bool], batch: int) -> int:
    """fp32 elements the gradient arena needs behind ``numel + TAIL`` for the per-m-tile bias-gradient partials and
    the loss partials."""
    n_mt = _cdiv(batch, 128)
    nb = sum(_pad(d[l], 64) for l in range(1, len(d)) if has_bias[l - 1])
    return 64 + (n_mt - 1) * nb + _pad(2 * n_mt, 64)


def _addr(t: torch.Tensor, elem_off: int = 0) -> int:
    return t.data_ptr() + elem_off * t.element_size()


@dataclasses.dataclass
class TmapSpec:
    tensor: torch.Tensor      # 2-D view [rows, cols] (unit inner stride)
    rows: int
    cols: int
    ld: int
    box_rows: int


class MegaPlan:
    """Task graph of one replica.  ``plan`` is a FusedMLP (device buffers, arenas, segment table)."""

    def __init__(self, plan, spec, patience: int, shard_lo: int, owned: Tuple[int, int], window: Optional[Tuple[int, int]],
                 x_src: torch.Tensor, tgt_src: torch.Tensor, flags: torch.Tensor, peers: Optional[Dict[str, List[int]]] = None,
                 grid: int = 148, chunk: int = 0):
        self.plan, self.spec, self.patience = plan, spec, patience
        self.world, self.rank = plan.world, plan.rank
        self.shard_lo, self.owned = shard_lo, owned
        self.window = window
        self.x_src, self.tgt_src = x_src, tgt_src
        self.flags = flags
        self.grid = grid
        self.peers = peers
        self.ok = False
        self.reason = ""
        self._build(chunk)

    # ------------------------------------------------------------------------------------------ build
    def _tmap(self, t: torch.Tensor, rows: int, cols: int, box_rows: int) -> int:
        key = (t.data_ptr(), rows, cols, t.stride(0), box_rows)
        i = self._tmap_index.get(key)
        if i is None:
            i = len(self.tmap_specs)
            self._tmap_index[key] = i
            self.tmap_specs.append(TmapSpec(t, rows, cols, t.stride(0), box_rows))
        return i

    def _build(self, chunk: int) -> None:
        plan = self.plan
        L, d, B = plan.L, plan.d, plan.B
        W, me = self.world, self.rank
        dev, adt = plan.device, plan.act_dtype
        relu = [ls.relu for ls in plan.lowered.layers]
        has_bias = [b is not None for b in plan.bias]
        n_mt = _cdiv(B, 128)
        self.n_mt = n_mt
        if 4 * L + 2 > MS_MAX_FLAGS or n_mt > 8 or not plan.fuse_loss:
            self.reason = "shape outside the K-STEP limits"
            return
        windowed = self.window is not None
        fu = plan.fu
        numel = fu.numel
        # ---- gradient-arena extras: bias partials of m-tiles >= 1, loss partials [2][n_mt]
        seg_rows = sorted(plan.segment_rows, key=lambda r: r[0])
        self.seg_rows = seg_rows
        seg_index = {r[0]: i for i, r in enumerate(seg_rows)}
        xoff = _pad(numel + TAIL, 64)
        boff, nb = {}, 0
        for l in range(1, L + 1):
            if has_bias[l - 1]:
                boff[l] = nb
                nb += _pad(d[l], 64)
        self.nb = nb
        loff = xoff + (n_mt - 1) * nb
        self.loss_off = loff
        if loff + 2 * n_mt > fu.grad.numel():
            self.reason = "gradient arena has no room for the K-STEP partial slots"
            return
        w_off_arena = [plan.arena_pos[id(ls.module.weight)] for ls in plan.lowered.layers]
        b_off_arena = [plan.arena_pos[id(ls.module.bias)] if ls.module.bias is not None else -1 for ls in plan.lowered.layers]
        # ---- double-buffered activations / dZ
        self.acts = [None] + [[torch.zeros(B, _pad(d[l]), dtype=adt, device=dev)[:, : d[l]] for _ in range(2)] for l in range(1, L)]
        self.dz = [None] + [[torch.zeros(B, _pad(d[l]), dtype=adt, device=dev)[:, : d[l]] for _ in range(2)] for l in range(1, L + 1)]
        self.tmap_specs: List[TmapSpec] = []
        self._tmap_index: Dict[tuple, int] = {}

        # ---- counters / flags
        def c_act(l): return l                      # noqa: E704  F_l tiles done (l = 1..L-1)
        def c_dz(l): return L + l                   # noqa: E704  dZ_l tiles done (l = 1..L)
        def c_wg(l): return 2 * L + l               # noqa: E704  wgrad tiles of W_l done
        def c_upd(s): return 3 * L + 1 + s          # noqa: E704  update chunks of segment s done (this rank)
        self.decide_cnt = 5 * L + 2
        self.n_counters = 5 * L + 3
        def f_dz(l): return l                       # noqa: E704  "dZ_l (hence db_l; l = L: the loss) complete on rank q", l = 1..L
        def f_gw(l): return L + l                   # noqa: E704  "dW_l complete on rank q"
        def f_pub(s): return 2 * L + 1 + s          # noqa: E704  "owner o published its part of segment s"
        all_mask = (1 << W) - 1
        EARLY = 16      # wait needed by the B operand (weights / resident batch): the producer loads B before the rest is met
        single = W == 1 and os.environ.get("SPARKTORCH_MEGA_FLAGS", "0") != "1"   # one GPU: plain counters, no flag hop

        def w_cnt(idx, need, lag=0, early=False):
            return [1 | (EARLY if early else 0), idx, need, lag]

        # cross-GPU dependencies: epoch flags raised by the last arriver of a group of tasks (kind 2), or - SPARKTORCH_MEGA_XCNT=1,
        # experimental - counters in the consumer's symmetric block that EVERY producer task bumps itself on every consumer rank
        # (kind 3; `arrivals` = increments per step from all ranks together).  Measured at 2 GPUs: 49.6 us/step against 41.7 with
        # flags - the system-scope fence every task then needs costs more than the last-arriver hop it saves.
        xcnt = W > 1 and os.environ.get("SPARKTORCH_MEGA_XCNT", "0") == "1"
        self.xcnt = xcnt

        def w_flag(idx, lag, mask, early=False, arrivals=0):
            if xcnt:
                assert arrivals > 0
                return [3 | (EARLY if early else 0), idx, arrivals, lag]
            return [2 | (EARLY if early else 0), idx, lag, mask]
        # ownership: flat ranges of float4s (ZeRO-1 layout of parallel/p2p.py)
        n4 = numel // 4
        per = _cdiv(n4, W) * 4

        def owners_of(lo: int, hi: int) -> int:
            m = 0
            for r in range(W):
                a, b = max(lo, per * r), min(hi, min(numel, per * (r + 1)))
                if a < b:
                    m |= 1 << r
            return m

        seg_of_w = [seg_index[o] for o in w_off_arena]
        seg_of_b = [seg_index[o] if o >= 0 else -1 for o in b_off_arena]
        seg_range = [(r[0], (seg_rows[i + 1][0] if i + 1 < len(seg_rows) else numel)) for i, r in enumerate(seg_rows)]
        seg_owner = [owners_of(a, b) for a, b in seg_range]

        ch = chunk or self._pick_chunk(self.owned[1] - self.owned[0])
        self.chunk = ch

        def my_chunks(sidx: int) -> int:
            lo, hi = max(seg_range[sidx][0], self.owned[0]), min(seg_range[sidx][1], self.owned[1])
            return _cdiv(hi - lo, ch) if lo < hi else 0

        def all_chunks(sidx: int) -> int:          # update chunks of segment sidx on all owners together
            n = 0
            for r in range(W):
                lo, hi = max(seg_range[sidx][0], min(numel, per * r)), min(seg_range[sidx][1], min(numel, per * (r + 1)))
                n += _cdiv(hi - lo, ch) if lo < hi else 0
            return n

        # one GPU: Adam/SGD is applied by the weight-gradient tiles themselves (no update phase for the weights)
        # (opt-in: with thread = weight row the in-place update writes 32-byte pieces of 32 different rows per instruction;
        #  the separate, fully coalesced update chunks are faster for the MNIST-sized layers - profiles/kstep_timeline_*.md)
        fused_opt = single and os.environ.get("SPARKTORCH_MEGA_FUSED_OPT", "0") == "1"
        self.fused_opt = fused_opt
        split = int(os.environ.get("SPARKTORCH_MEGA_EPI_SPLIT", "2"))       # CTAs per 64-wide backward tile (1, 2 or 4)
        if split not in (1, 2, 4):
            split = 2
        # chained head: the CTAs of the last layer's dgrad tiles recompute the head (logits, loss, dZ_L) of their rows and feed
        # dZ_L to the dgrad MMA through shared memory; the separate head phase (and its signal/wait hop) disappears
        chain = (os.environ.get("SPARKTORCH_MEGA_CHAIN", "1") == "1" and L >= 2 and d[L] <= 32 and not fused_opt)
        self.chain = chain
        # first layer: the batch (A) has no dependency inside the step, the weights (B) have - load A while waiting for them
        early_a = os.environ.get("SPARKTORCH_MEGA_EARLY_A", "1") == "1"
        wg_tiles = {l: _cdiv(d[l], 128) * sum(len([c for c in range(0, 64, 64 // split) if d[l - 1] - nt * 64 - c > 0])
                                              for nt in range(_cdiv(d[l - 1], 64))) for l in range(1, L + 1)}

        def w_pub(sidx: int, early=True):          # "segment sidx as of the previous step is published everywhere"
            if fused_opt and sidx in seg_of_w:
                l = seg_of_w.index(sidx) + 1
                return w_cnt(c_wg(l), wg_tiles[l], lag=1, early=early)
            if single:
                return w_cnt(c_upd(sidx), my_chunks(sidx), lag=1, early=early)
            return w_flag(f_pub(sidx), 1, seg_owner[sidx], early=early, arrivals=all_chunks(sidx))

        tasks: List[dict] = []

        def new_task(**kw) -> dict:
            t = dict(type=0, flags=0, tmap_a=[0, 0], tmap_b=[0, 0], m0=0, n0=0, M=0, N=0, num_kb=0, block_n=0, ld_mask=0,
                     ld_out=0, ld_f32=0, ld_target=0, bias=0, mask=[0, 0], out_bf16=[0, 0], out_f32=0, colsum=0, labels=0,
                     loss_slot=[0, 0], loss_scale=0.0, seg=0, e0=0, len=0, n_part=1, part_off=0, part_stride=0,
                     wait=[], sig=[-1, 0, -1, 0], ep_c0=0, ep_cols=0, cost=1.0, group="", after_prev=False)
            t.update(kw)
            tasks.append(t)
            return t

        def bn_fwd(l: int) -> int:
            if l == L:
                return 16 if d[L] <= 16 else 32
            n = d[l]
            if n <= 16:
                return 16
            return 32 if n_mt * _cdiv(n, 32) <= 64 else 64

        def src_of(l: int, par: int) -> torch.Tensor:      # A_{l} (l = 0: the batch)
            return self.x_src if l == 0 else self.acts[l][par]

        def colsum_ptr(l: int, mt: int) -> int:
            if mt == 0:
                return _addr(fu.grad, b_off_arena[l - 1])
            return _addr(fu.grad, xoff + (mt - 1) * nb + boff[l])

        n_f_tiles = [0] * (L + 1)
        head_of: Dict[int, dict] = {}
        # ---------------------------------------------------------------- forward
        for l in range(1, L + 1):
            bn = bn_fwd(l)
            Wb = plan.Wb[l - 1]
            K = d[l - 1]
            nkb = _cdiv(K, 64)
            ntl = _cdiv(d[l], bn)
            n_f_tiles[l] = n_mt * ntl
            for mt in range(n_mt):
                for nt in range(ntl):
                    fl = (F_BIAS if has_bias[l - 1] else 0) | (F_RELU if relu[l - 1] else 0)
                    src_rows = self.x_src.shape[0] if l == 1 else B
                    if l == L and chain:
                        assert ntl == 1
                    t = new_task(type=MS_GEMM, m0=mt * 128, n0=nt * bn, M=B, N=d[l], num_kb=nkb, block_n=bn,
                                 tmap_a=[self._tmap(src_of(l - 1, p), src_rows, K, 128) for p in range(2)],
                                 tmap_b=[self._tmap(Wb, d[l], K, bn)] * 2,
                                 bias=_addr(plan.bias[l - 1]) if has_bias[l - 1] else 0, cost=1.2 + 0.15 * nkb, group=f"F{l}")
                    if l == 1 and windowed:
                        fl |= F_A_WIN
                    if l == 1 and early_a:
                        fl |= F_EARLY_A
                    w = []
                    if l > 1:
                        w.append(w_cnt(c_act(l - 1), n_f_tiles[l - 1]))
                    b_early = not (l == 1 and early_a)      # early-A tiles: the weights are the late operand
                    w.append(w_pub(seg_of_w[l - 1], early=b_early))
                    if has_bias[l - 1]:
                        w.append(w_pub(seg_of_b[l - 1], early=b_early))
                    t["wait"] = w
                    if l < L:
                        fl |= F_OUT_BF16
                        t["out_bf16"] = [_addr(self.acts[l][p]) for p in range(2)]
                        t["ld_out"] = self.acts[l][0].stride(0)
                        t["sig"] = [c_act(l), n_f_tiles[l], -1, 0]
                    else:
                        fl |= F_OUT_BF16 | (F_LOSS_CE if plan.kind == "ce" else F_LOSS_MSE) | (F_LBL_WIN if windowed else 0)
                        t["out_bf16"] = [_addr(self.dz[L][p]) for p in range(2)]
                        t["ld_out"] = self.dz[L][0].stride(0)
                        t["labels"] = _addr(self.tgt_src)
                        t["ld_target"] = self.tgt_src.stride(0) if self.tgt_src.dim() > 1 else 0
                        t["loss_slot"] = [_addr(fu.grad, loff + p * n_mt + mt) for p in range(2)]
                        t["loss_scale"] = (1.0 / B) if plan.kind == "ce" else 1.0 / (B * d[L])
                        if has_bias[L - 1]:
                            fl |= F_COLSUM
                            t["colsum"] = colsum_ptr(L, mt)
                        # the head's tiles complete dZ_L, db_L and the loss of this rank: flag gr(b_L) doubles as "loss ready"
                        t["sig"] = [c_dz(L), n_f_tiles[L], -1 if single else f_dz(L), all_mask]
                    t["flags"] = fl
                    if l == L and chain:        # also the template of the (quiet) head copies in front of the dgrad tiles of m-tile mt
                        head_of[mt] = t
        self.loss_flag = f_dz(L)
        # ---------------------------------------------------------------- decision
        def decide_task():
            new_task(type=MS_DECIDE, wait=[w_cnt(c_dz(L), n_f_tiles[L]) if single else w_flag(self.loss_flag, 0, all_mask, arrivals=W * n_f_tiles[L])],
                     cost=0.5, group="D")

        decide_task()
        # ---------------------------------------------------------------- backward, last layer first
        n_dz_tiles = [0] * (L + 1)
        n_dz_tiles[L] = n_f_tiles[L]
        def update_tasks(s: int, gr_waits: List[list], n_part: int, part_off: int) -> None:
            lo, hi = max(seg_range[s][0], self.owned[0]), min(seg_range[s][1], self.owned[1])
            if lo >= hi:
                return
            starts = list(range(lo, hi, ch))
            for a in starts:
                new_task(type=MS_UPDATE, seg=s, e0=a, len=min(ch, hi - a), n_part=n_part, part_off=part_off, part_stride=nb,
                         wait=gr_waits, sig=[c_upd(s), len(starts), -1 if single else f_pub(s), all_mask],
                         cost=1.5, group=f"U{s}")

        def w_dz(l):       # "dZ_l (and db_l) of this step complete on every replica"
            return w_cnt(c_dz(l), n_dz_tiles[l]) if single else w_flag(f_dz(l), 0, all_mask, arrivals=W * n_dz_tiles[l])

        # the 64-wide backward tile is run by `split` CTAs that drain one window of its columns each (epilogue time ~ columns per thread)
        wins = lambda nt_, N_: [(c, min(64 // split, N_ - nt_ * 64 - c)) for c in range(0, 64, 64 // split) if N_ - nt_ * 64 - c > 0]   # noqa: E731
        for l in range(L, 0, -1):
            chained = chain and l == L
            if l > 1:
                n_dz_tiles[l - 1] = n_mt * sum(len(wins(nt, d[l - 1])) for nt in range(_cdiv(d[l - 1], 64)))
            def bias_update():      # db_l is complete once dZ_l is (column sums of its producers)
                # chained head: the quiet copies of the head read b_L until the last dgrad tile of the layer is done
                update_tasks(seg_of_b[l - 1], [w_dz(l)] + ([w_dz(l - 1)] if chained else []), n_mt, xoff + boff[l] - b_off_arena[l - 1])

            if has_bias[l - 1] and not chained:
                bias_update()
            if l > 1:
                # dgrad: dZ_{l-1}[B, d_{l-1}] = dZ_l[B, d_l] W_l[d_l, d_{l-1}]  (B operand MN-major, 64-wide tiles)
                K = d[l]
                nkb = _cdiv(K, 64)
                ntl = _cdiv(d[l - 1], 64)
                for mt in range(n_mt):
                    for j, (nt, (wc0, wcn)) in enumerate([(nt_, w_) for nt_ in range(ntl) for w_ in wins(nt_, d[l - 1])]):
                        fl = F_B_MN | F_OUT_BF16
                        if chained:
                            # the head copy that feeds this tile keeps everything on chip; dZ_L / db_L / the loss partial for the
                            # weight gradient, the bias update and the decision come from the regular head tiles of the forward
                            # pass (their stores and signal are then off the critical path of this tile)
                            h = dict(head_of[mt])
                            h["flags"] |= F_CHAIN_HEAD | F_QUIET
                            h["sig"] = [-1, 0, -1, 0]
                            h["group"] = f"H{l}"
                            tasks.append(h)
                            fl |= F_CHAIN_TAIL
                        t = new_task(type=MS_GEMM, m0=mt * 128, n0=nt * 64, M=B, N=d[l - 1], num_kb=nkb, block_n=64,
                                     tmap_a=[self._tmap(self.dz[l][p], B, K, 128) for p in range(2)],
                                     tmap_b=[self._tmap(plan.Wb[l - 1], K, d[l - 1], 64)] * 2,
                                     out_bf16=[_addr(self.dz[l - 1][p]) for p in range(2)], ld_out=self.dz[l - 1][0].stride(0),
                                     ep_c0=wc0, ep_cols=wcn if split > 1 else 0,
                                     wait=([] if chained else [w_cnt(c_dz(l), n_dz_tiles[l])]) + [w_pub(seg_of_w[l - 1])],
                                     cost=(0.9 if chained else 1.2) + 0.15 * nkb, group=f"B{l}", after_prev=chained)
                        if relu[l - 2]:
                            fl |= F_MASK
                            t["mask"] = [_addr(self.acts[l - 1][p]) for p in range(2)]
                            t["ld_mask"] = self.acts[l - 1][0].stride(0)
                        if has_bias[l - 2]:
                            fl |= F_COLSUM
                            t["colsum"] = colsum_ptr(l - 1, mt)
                        t["sig"] = [c_dz(l - 1), n_dz_tiles[l - 1], -1 if single else f_dz(l - 1), all_mask]
                        t["flags"] = fl
            if has_bias[l - 1] and chained:       # (after the tiles it waits for: every CTA's list stays a topological order)
                bias_update()
            # wgrad: dW_l[d_l, d_{l-1}] = dZ_l^T A_{l-1}   (both operands MN-major, reduction over the batch)
            nkb = _cdiv(B, 64)
            mtl, ntl = _cdiv(d[l], 128), _cdiv(d[l - 1], 64)
            sw = seg_of_w[l - 1]
            gwins = lambda nt_: [(c, min(64 // split, d[l - 1] - nt_ * 64 - c)) for c in range(0, 64, 64 // split) if d[l - 1] - nt_ * 64 - c > 0]   # noqa: E731
            n_g = mtl * sum(len(gwins(nt)) for nt in range(ntl))
            wg_tiles[l] = n_g
            for mt in range(mtl):
                for nt, (wc0, wcn) in [(nt_, w_) for nt_ in range(ntl) for w_ in gwins(nt_)]:
                    src_rows = self.x_src.shape[0] if l == 1 else B
                    new_task(type=MS_GEMM, flags=F_A_MN | F_B_MN | F_OUT_F32 | (F_B_WIN if (l == 1 and windowed) else 0) | (F_FUSED_OPT if fused_opt else 0),
                             seg=sw, m0=mt * 128, n0=nt * 64, M=d[l], N=d[l - 1], num_kb=nkb, block_n=64,
                             tmap_a=[self._tmap(self.dz[l][p], B, d[l], 64) for p in range(2)],
                             tmap_b=[self._tmap(src_of(l - 1, p), src_rows, d[l - 1], 64) for p in range(2)],
                             out_f32=_addr(fu.grad, w_off_arena[l - 1]), ld_f32=d[l - 1], ep_c0=wc0, ep_cols=wcn if split > 1 else 0,
                             wait=[w_cnt(c_dz(l), n_dz_tiles[l])] + ([w_cnt(c_act(l - 1), n_f_tiles[l - 1], early=True)] if l > 1 else [])
                             # fused optimizer: this tile overwrites W_l, which this step's dgrad B_l still reads until dZ_{l-1} is complete
                             + ([w_cnt(c_dz(l - 1), n_dz_tiles[l - 1])] if (fused_opt and l > 1) else []),
                             sig=[c_wg(l), n_g, -1 if single else f_gw(l), all_mask],
                             cost=1.5 + 0.15 * nkb, group=f"G{l}")
            # W_l may only be overwritten (on EVERY replica: the owner publishes everywhere) once every replica's dgrad of
            # this step has read it: dZ_{l-1} complete  <=>  the B_l tiles are done
            if not fused_opt:
                w_gw = w_cnt(c_wg(l), n_g) if single else w_flag(f_gw(l), 0, all_mask, arrivals=W * n_g)
                update_tasks(sw, [w_gw] + ([w_dz(l - 1)] if l > 1 else []), 1, 0)
        self.task_dicts = tasks
        if not self._place():
            return
        self.ok = True

    def _pick_chunk(self, owned_elems: int) -> int:
        # 256 update threads x two float4 in flight = 2048 floats per pass; one pass per chunk whenever the owned shard
        # spreads over the grid that way
        c = _pad(_cdiv(max(owned_elems, 1), max(1, self.grid)), 512)
        return int(min(2048, max(512, c)))

    # ------------------------------------------------------------------------------------------ placement
    def _place(self) -> bool:
        """Earliest-start list scheduling of one step on ``grid`` CTAs with a rough cost model (micro-seconds); the task
        order of every CTA is a topological order of the step, which is what makes the dataflow dead-lock free."""
        G = self.grid
        free = [0.0] * G
        cnt_done: Dict[int, float] = {}
        flag_done: Dict[int, float] = {}
        lists: List[List[int]] = [[] for _ in range(G)]
        prev_cta = 0
        for ti, t in enumerate(self.task_dicts):
            ready = 0.0
            for k, idx, need, m in t["wait"]:
                if (k & 15) == 1 and m == 0:
                    ready = max(ready, cnt_done.get(idx, 0.0))
                elif (k & 15) == 2 and need == 0:
                    ready = max(ready, flag_done.get(idx, 0.0) + (1.5 if self.world > 1 else 0.0))
                elif (k & 15) == 3 and m == 0:
                    ready = max(ready, flag_done.get(idx, 0.0) + 1.0)
            if t.get("after_prev"):
                best = prev_cta                 # second half of a chained pair: same CTA, right behind the head copy
            elif t["type"] == MS_DECIDE:
                # on the CTA of the first forward tile: everything of step s+1 then follows the decision of step s in
                # program order, which is what keeps the loss partial slots (double buffered) safe to reuse
                best = next(c for c in range(G) if lists[c] and lists[c][0] == 0)
            else:
                # earliest start; among equals the CTA with the fewest tasks so far (an idle SM beats queueing behind a finished
                # task: epilogues are latency bound, a second tile on the same SM would wait for the first one's stores)
                best = min(range(G), key=lambda c: (max(ready, free[c]), len(lists[c]), c))
            start = max(ready, free[best])
            fin = start + t["cost"]
            free[best] = fin
            lists[best].append(ti)
            prev_cta = best
            sc, _tot, sf, _sm = t["sig"]
            if sc >= 0:
                cnt_done[sc] = max(cnt_done.get(sc, 0.0), fin)
            if sf >= 0:
                flag_done[sf] = max(flag_done.get(sf, 0.0), fin)
        self.est_step_us = max(free)
        if max(len(x) for x in lists) > MS_MAX_TASKS:
            self.reason = "more tasks per CTA than the kernel's descriptor table holds"
            return False
        used = [x for x in lists if x]
        self.grid_used = len(used)
        arr = np.zeros((self.grid_used, MS_MAX_TASKS), dtype=MS_TASK)
        n_tasks = np.zeros(self.grid_used, dtype=np.int32)
        self.cta_of_task = {}
        for c, lst in enumerate(used):
            n_tasks[c] = len(lst)
            for j, ti in enumerate(lst):
                self.cta_of_task[ti] = (c, j)
                t = self.task_dicts[ti]
                r = arr[c, j]
                for k in ("type", "flags", "m0", "n0", "M", "N", "num_kb", "block_n", "ld_mask", "ld_out", "ld_f32", "ld_target",
                          "bias", "out_f32", "colsum", "labels", "loss_scale", "seg", "e0", "len", "n_part", "part_off", "part_stride", "ep_c0", "ep_cols"):
                    r[k] = t[k]
                for k in ("tmap_a", "tmap_b", "mask", "out_bf16", "loss_slot"):
                    r[k] = t[k]
                w = np.zeros((3, 4), dtype=np.int32)
                assert len(t["wait"]) <= 3
                for i, e in enumerate(t["wait"]):
                    w[i] = e
                r["wait"] = w
                r["sig"] = t["sig"]
        self.tasks, self.n_tasks = arr, n_tasks
        return True

    # ------------------------------------------------------------------------------------------ device objects
    def materialize(self):
        """Create the native ``MlpStep`` (GPU only): tensor maps, task table, peer table, counters."""
        from sparktorch_b200.ops import native

        C = native.load(required=True)
        assert C.ms_task_bytes() == MS_TASK.itemsize and C.ms_max_tasks() == MS_MAX_TASKS
        plan, fu, dev = self.plan, self.plan.fu, self.plan.device
        maps = b"".join(C.make_tmap_bytes(_addr(s.tensor), s.rows, s.cols, s.ld, s.box_rows) for s in self.tmap_specs)
        self.d_tmaps = torch.frombuffer(bytearray(maps), dtype=torch.uint8).to(dev)
        self.d_tasks = torch.from_numpy(self.tasks.view(np.uint8).reshape(-1)).to(dev)
        self.d_ntasks = torch.from_numpy(self.n_tasks).to(dev)
        W = self.world
        pt = self.peers or {"flags": [_addr(self.flags)], "grad": [_addr(fu.grad)], "param": [_addr(fu.param)], "wbf16": [_addr(fu.wbf16)]}
        table = []
        for k in ("flags", "grad", "param", "wbf16"):
            table += list(pt[k]) + [0] * (8 - len(pt[k]))
        self.d_peers = torch.tensor(table, dtype=torch.int64).view(torch.uint8).to(dev)
        self.d_counters = torch.zeros(max(64, self.n_counters) * CNT_STRIDE, dtype=torch.int32, device=dev)
        self.d_error = torch.zeros(4, dtype=torch.int32, device=dev)
        mc = pt.get("mc", {}) if W > 1 else {}
        s = self.spec
        self.native = C.MlpStep(
            self.d_tasks, self.d_ntasks, self.d_tmaps, self.d_peers, self.d_counters, _addr(self.flags), W, self.rank,
            _addr(fu.grad), _addr(fu.param), _addr(fu.wbf16), mc.get("grad", 0), mc.get("param", 0), mc.get("wbf16", 0),
            fu.m, fu.v, self.shard_lo, fu._segs, fu.state, fu.loss_hist, fu.mailbox.data_ptr(), s.kind_id, s.lr, s.beta1,
            s.beta2, s.eps, s.weight_decay, s.momentum, s.dampening, s.nesterov, s.maximize, self.patience, self.decide_cnt,
            self.n_mt, self.loss_off, self.window[1] if self.window else 0, self.window[0] if self.window else 1, self.d_error)
        if self.xcnt:
            self.native.set_xcnt(int(mc.get("flags", 0)))
        if os.environ.get("SPARKTORCH_MEGA_SYSFENCE", "0") == "1":      # conservative: system-scope fence in front of every flag
            self.native.set_sys_fence_all(True)
        return self.native

    def describe(self) -> dict:
        groups: Dict[str, int] = {}
        for t in self.task_dicts:
            groups[t["group"]] = groups.get(t["group"], 0) + 1
        return {"ctas": self.grid_used, "tasks_per_step": len(self.task_dicts), "groups": groups, "tensor_maps": len(self.tmap_specs),
                "max_tasks_per_cta": int(self.n_tasks.max()), "update_chunk": self.chunk, "est_step_us": round(self.est_step_us, 2)}


# =================================================================================================== CPU emulation
class _Registry:
    def __init__(self):
        self.items: List[Tuple[int, int, torch.Tensor]] = []

    def add(self, t: torch.Tensor) -> None:
        base = t
        while base._base is not None:
            base = base._base
        flat = base.reshape(-1) if base.is_contiguous() else base
        lo = flat.data_ptr()
        self.items.append((lo, lo + flat.numel() * flat.element_size(), flat))

    def resolve(self, addr: int) -> Tuple[torch.Tensor, int]:
        for lo, hi, flat in self.items:
            if lo <= addr < hi:
                assert (addr - lo) % flat.element_size() == 0
                return flat, (addr - lo) // flat.element_size()
        raise KeyError(hex(addr))

    def view2d(self, addr: int, rows: int, cols: int, ld: int) -> torch.Tensor:
        flat, off = self.resolve(addr)
        return flat.as_strided((rows, cols), (ld, 1), off)

    def view1d(self, addr: int, n: int) -> torch.Tensor:
        flat, off = self.resolve(addr)
        return flat[off: off + n]


class MegaEmulator:
    """Interprets the task tables of W replicas (one ``MegaPlan`` each, all in this process on CPU tensors) with the
    kernel's semantics: every CTA walks its list in order, a task runs when its waits are satisfied, tasks of different
    CTAs / ranks interleave in a random order.  Raises on dead-lock."""

    def __init__(self, plans: List[MegaPlan]):
        self.plans = plans
        self.W = len(plans)
        self.reg = _Registry()
        for mp in plans:
            fu = mp.plan.fu
            for t in (fu.grad, fu.param, fu.wbf16, mp.x_src, mp.tgt_src, mp.flags):
                self.reg.add(t)
            for l in range(1, mp.plan.L):
                for p in range(2):
                    self.reg.add(mp.acts[l][p])
            for l in range(1, mp.plan.L + 1):
                for p in range(2):
                    self.reg.add(mp.dz[l][p])
        self.counters = [np.zeros(mp.n_counters, dtype=np.int64) for mp in plans]
        self.state = [dict(stopped=False, es_seeded=False, es_best=0.0, es_bad=0, step=0, iter=0, losses=[]) for _ in plans]
        self.chain_smem: Dict[Tuple[int, int], torch.Tensor] = {}

    def _tile_operand(self, spec: TmapSpec, mn: bool, r0: int, k_off: int, width: int, kp: int) -> torch.Tensor:
        """[width, kp] fp32 operand tile with zero fill outside the tensor (what TMA delivers)."""
        src = spec.tensor
        out = torch.zeros(width, kp, dtype=torch.float32)
        if mn:      # stored [K rows, MN contiguous]
            ks, ke = k_off, min(spec.rows, k_off + kp)
            cs, ce = r0, min(spec.cols, r0 + width)
            if ks < ke and cs < ce:
                out[: ce - cs, : ke - ks] = src[ks:ke, cs:ce].to(torch.float32).t()
        else:
            rs, re = r0, min(spec.rows, r0 + width)
            ce = min(spec.cols, kp)
            if rs < re:
                out[: re - rs, :ce] = src[rs:re, :ce].to(torch.float32)
        return out

    def _waits_ok(self, ri: int, t, epoch: int) -> bool:
        mp = self.plans[ri]
        for k, idx, need, mask in t["wait"]:
            k &= 15
            if k == 1 and self.counters[ri][idx] < (epoch - mask) * need:
                return False
            if k == 2:
                fl = mp.flags
                for r in range(self.W):
                    if (mask >> r) & 1 and int(fl[idx * 8 + r]) < epoch - need:
                        return False
            if k == 3 and int(mp.flags[XCNT_BASE + idx * CNT_STRIDE]) < (epoch - mask) * need:
                return False
        needs_decision = int(t["type"]) == MS_UPDATE or (int(t["type"]) == MS_GEMM and int(t["flags"]) & F_FUSED_OPT)
        if needs_decision and mp.patience > 0 and (int(self.counters[ri][mp.decide_cnt]) >> 1) < epoch:
            return False
        return True

    def _signal(self, ri: int, t, epoch: int) -> None:
        sc, tot, sf, sm = (int(x) for x in t["sig"])
        if sc < 0:
            return
        self.counters[ri][sc] += 1
        if sf >= 0 and self.plans[ri].xcnt:
            for r in range(self.W):
                if (sm >> r) & 1:
                    self.plans[r].flags[XCNT_BASE + sf * CNT_STRIDE] += 1
            return
        if sf >= 0 and self.counters[ri][sc] == epoch * tot:
            for r in range(self.W):
                if (sm >> r) & 1:
                    self.plans[r].flags[sf * 8 + ri] = epoch

    def _fused_opt_tile(self, ri: int, t, grad, m0, n0, rows, cols, step1) -> None:
        """one GPU: the weight-gradient tile applies the optimizer to the parameters it covers"""
        from sparktorch_b200.runtime.optim import apply_update_

        mp = self.plans[ri]
        fu = mp.plan.fu
        flat, off = self.reg.resolve(int(t["out_f32"]))
        assert flat.data_ptr() == fu.grad.data_ptr()
        ld = int(t["ld_f32"])
        idx = (off + (torch.arange(m0, m0 + rows)[:, None] * ld + torch.arange(n0, n0 + cols)[None, :])).reshape(-1)
        p = fu.param[idx].clone()
        st = {}
        if fu.m is not None:
            st["m" if mp.spec.kind != "sgd" else "buf"] = fu.m[idx - mp.shard_lo].clone()
        if fu.v is not None:
            st["v"] = fu.v[idx - mp.shard_lo].clone()
        apply_update_(mp.spec, p, grad.reshape(-1).clone(), st, step1)
        fu.param[idx] = p
        if fu.m is not None:
            fu.m[idx - mp.shard_lo] = st["m" if mp.spec.kind != "sgd" else "buf"]
        if fu.v is not None:
            fu.v[idx - mp.shard_lo] = st["v"]
        s_off, _r, s_cols, w_off, w_ld = (int(x) for x in mp.seg_rows[int(t["seg"])][:5])
        assert s_off == off and s_cols == ld
        if w_off >= 0:
            widx = (w_off + (torch.arange(m0, m0 + rows)[:, None] * w_ld + torch.arange(n0, n0 + cols)[None, :])).reshape(-1)
            fu.wbf16[widx] = p.to(fu.wbf16.dtype)

    def _gemm(self, ri: int, t, par: int, win_off: int, step1: int = 0, cta: int = 0) -> None:
        mp = self.plans[ri]
        fl, bn, kp = int(t["flags"]), int(t["block_n"]), int(t["num_kb"]) * 64
        m0, n0, M, N = int(t["m0"]), int(t["n0"]), int(t["M"]), int(t["N"])
        sa, sb = mp.tmap_specs[int(t["tmap_a"][par])], mp.tmap_specs[int(t["tmap_b"][par])]
        quiet = bool(fl & F_QUIET)
        if fl & F_CHAIN_TAIL:                               # A = what the head copy before it left in the CTA's shared memory
            A = self.chain_smem.pop((ri, cta))[:, :kp]
        elif fl & F_A_MN:
            A = self._tile_operand(sa, True, m0, 0, 128, kp)
        else:
            A = self._tile_operand(sa, False, m0 + (win_off if (fl & F_A_WIN) else 0), 0, 128, kp)
        Bm = self._tile_operand(sb, bool(fl & F_B_MN), n0, win_off if (fl & F_B_WIN) else 0, bn, kp)
        acc = A @ Bm.t()                                    # [128, bn]
        rows = max(0, min(128, M - m0))
        w0 = int(t["ep_c0"])
        wc = int(t["ep_cols"]) or bn                       # epilogue window: this task drains columns [w0, w0 + wc) of the tile
        n0 = n0 + w0
        cols = max(0, min(wc, N - n0))
        acc = acc[:rows, w0: w0 + cols].clone()
        if fl & F_BIAS:
            acc += self.reg.view1d(int(t["bias"]), N)[n0: n0 + cols].to(torch.float32)
        if fl & F_RELU:
            acc.clamp_(min=0)
        if fl & F_MASK:
            mk = self.reg.view2d(int(t["mask"][par]), M, N, int(t["ld_mask"]))[m0: m0 + rows, n0: n0 + cols]
            acc = torch.where(mk.to(torch.float32) > 0, acc, torch.zeros_like(acc))
        lbl_off = win_off if (fl & F_LBL_WIN) else 0
        if fl & F_LOSS_CE:
            flat, off = self.reg.resolve(int(t["labels"]))
            y = flat[off + m0 + lbl_off: off + m0 + lbl_off + rows].long()
            lse = torch.logsumexp(acc, dim=1)
            li = (lse - acc.gather(1, y[:, None])[:, 0]).sum() * float(t["loss_scale"])
            sm = torch.softmax(acc, dim=1)
            sm[torch.arange(rows), y] -= 1.0
            acc = sm * float(t["loss_scale"])
            if not quiet:
                self.reg.view1d(int(t["loss_slot"][par]), 1)[0] = li
        elif fl & F_LOSS_MSE:
            tgt = self.reg.view2d(int(t["labels"]), m0 + lbl_off + rows, N, int(t["ld_target"]))[m0 + lbl_off:, :cols]
            dlt = acc - tgt.to(torch.float32)
            if not quiet:
                self.reg.view1d(int(t["loss_slot"][par]), 1)[0] = (dlt * dlt).sum() * float(t["loss_scale"])
            acc = 2.0 * dlt * float(t["loss_scale"])
        if fl & F_CHAIN_HEAD:
            buf = torch.zeros(128, 64, dtype=torch.float32)
            buf[:rows, :cols] = acc.to(mp.plan.act_dtype).to(torch.float32)
            self.chain_smem[(ri, cta)] = buf
        if quiet:
            return
        if fl & F_OUT_BF16:
            out = self.reg.view2d(int(t["out_bf16"][par]), M, N, int(t["ld_out"]))
            out[m0: m0 + rows, n0: n0 + cols] = acc.to(out.dtype)
        if fl & F_FUSED_OPT:
            if not (mp.patience > 0 and int(self.counters[ri][mp.decide_cnt]) & 1):
                self._fused_opt_tile(ri, t, acc, m0, n0, rows, cols, step1)
        elif fl & F_OUT_F32:
            out = self.reg.view2d(int(t["out_f32"]), M, N, int(t["ld_f32"]))
            out[m0: m0 + rows, n0: n0 + cols] = acc
        if fl & F_COLSUM:
            self.reg.view1d(int(t["colsum"]), N)[n0: n0 + cols] = acc.sum(dim=0)

    def _update(self, ri: int, t, step1: int) -> None:
        from sparktorch_b200.runtime.optim import apply_update_

        mp = self.plans[ri]
        if int(self.counters[ri][mp.decide_cnt]) & 1 and mp.patience > 0:
            return
        e0, n = int(t["e0"]), int(t["len"])
        g = torch.zeros(n, dtype=torch.float32)
        for q in self.plans:
            g += q.plan.fu.grad[e0: e0 + n]
            for pr in range(1, int(t["n_part"])):
                o = int(t["part_off"]) + (pr - 1) * int(t["part_stride"]) + e0
                g += q.plan.fu.grad[o: o + n]
        g /= self.W
        fu = mp.plan.fu
        p = fu.param[e0: e0 + n].clone()
        st = {}
        lo = e0 - mp.shard_lo
        if fu.m is not None:
            st["m" if mp.spec.kind != "sgd" else "buf"] = fu.m[lo: lo + n]
        if fu.v is not None:
            st["v"] = fu.v[lo: lo + n]
        apply_update_(mp.spec, p, g, st, step1)
        off, rows, cols, w_off, w_ld = (int(x) for x in mp.seg_rows[int(t["seg"])][:5])
        for q in self.plans:
            q.plan.fu.param[e0: e0 + n] = p
            if w_off >= 0:
                for j in range(n):
                    li = e0 + j - off
                    if li < rows * cols:
                        q.plan.fu.wbf16[w_off + (li // cols) * w_ld + li % cols] = p[j].to(q.plan.fu.wbf16.dtype)

    def _decide(self, ri: int, epoch: int, it: int, par: int) -> None:
        mp, st = self.plans[ri], self.state[ri]
        loss = 0.0
        for q in self.plans:
            for k in range(mp.n_mt):
                loss += float(q.plan.fu.grad[mp.loss_off + par * mp.n_mt + k])
        loss /= self.W
        stop_bit = 0
        if st["stopped"]:
            stop_bit = 1
        else:
            stop_now = False
            if mp.patience > 0:
                if not st["es_seeded"]:
                    st.update(es_seeded=True, es_best=loss, es_bad=0)
                elif math.isnan(loss):
                    stop_now = True
                elif loss < st["es_best"]:
                    st.update(es_best=loss, es_bad=0)
                else:
                    st["es_bad"] += 1
                    stop_now = st["es_bad"] >= mp.patience
            st["losses"].append(loss)
            if stop_now:
                st["stopped"] = True
                stop_bit = 1
            else:
                st["step"] += 1
            st["iter"] = it + 1
        if mp.patience > 0:
            self.counters[ri][mp.decide_cnt] = epoch * 2 + stop_bit

    def run(self, n_steps: int, iter0: int = 0, step0: int = 0, epoch0: int = 0, seed: int = 0) -> None:
        rng = np.random.default_rng(seed)
        pos = {(ri, c): [0, 0] for ri, mp in enumerate(self.plans) for c in range(mp.grid_used)}   # [step, task index]
        live = set(pos)
        while live:
            keys = list(live)
            rng.shuffle(keys)
            progressed = False
            for key in keys:
                ri, c = key
                mp = self.plans[ri]
                s, j = pos[key]
                if j >= int(mp.n_tasks[c]):
                    s, j = s + 1, 0
                    pos[key] = [s, j]
                if s >= n_steps:
                    live.discard(key)
                    progressed = True
                    continue
                t = mp.tasks[c, j]
                epoch = epoch0 + s + 1
                td = dict(wait=[tuple(int(x) for x in w) for w in t["wait"]], type=int(t["type"]), flags=int(t["flags"]))
                if not self._waits_ok(ri, td, epoch):
                    continue
                par = (epoch0 + s) & 1
                it = iter0 + s
                win_off = (it % mp.window[0]) * mp.window[1] if mp.window else 0
                ty = int(t["type"])
                if ty == MS_GEMM:
                    self._gemm(ri, t, par, win_off, step0 + s + 1, c)
                elif ty == MS_UPDATE:
                    self._update(ri, t, step0 + s + 1)
                elif ty == MS_DECIDE:
                    self._decide(ri, epoch, it, par)
                self._signal(ri, t, epoch)
                pos[key] = [s, j + 1]
                progressed = True
                if rng.random() < 0.5:
                    break                      # re-shuffle often: finer interleavings
            if not progressed:
                stuck = {k: pos[k] for k in live}
                raise RuntimeError(f"K-STEP task graph dead-locked: {stuck}")
