"""Fused MLP training / inference on B200: no autograd, no Python in the loop.

A model that lowers to ``Linear (+ReLU) ... Linear [+log_softmax]`` with
CrossEntropy/NLL or MSE loss and Adam/AdamW/SGD becomes a static step plan of
hand-written sm_100a kernels, captured once into a CUDA graph and replayed by
the native ``GraphRunner``:

    gather_batch                    device-side sampling of the step's rows
    L x gemm_bf16 (fwd)             tcgen05 GEMM, bias(+ReLU) fused, bf16 out
    softmax_ce | mse_loss           loss + dlogits + last bias grad, one pass
    L x gemm_bf16 (wgrad)           dW = dZ^T A straight into the fp32 gradient arena
                                    (both operands consumed MN-major: no transposes)
    (L-1) x gemm_bf16 (dgrad)       dZ_prev = (dZ W) * (A>0), bias grad by column sums
    allreduce_update (K-AR)         NVLink reduce-scatter + Adam/SGD + bf16 publish
                                    + early-stop decision + loss history, one kernel

This is the replacement of the reference hot loop (distributed.py:140-203) for
its flagship workloads (examples/simple_dnn.py); anything that does not lower
falls back to the eager engine.  bf16 tensor-core compute, fp32 master weights,
fp32 optimizer state (ZeRO-1 sharded across the data-parallel ranks).
"""
from __future__ import annotations

import dataclasses
import os
import time
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from sparktorch_b200.ops import native
from sparktorch_b200.runtime.flat import FlatArena
from sparktorch_b200.runtime.optim import parse_optimizer


def _pad(n: int, m: int = 8) -> int:
    return (n + m - 1) // m * m


@dataclasses.dataclass
class LinearSpec:
    module: nn.Linear
    relu: bool


@dataclasses.dataclass
class Lowered:
    layers: List[LinearSpec]
    final_log_softmax: bool


# ----------------------------------------------------------------------------- lowering
def lower_mlp(model: nn.Module) -> Optional[Lowered]:
    """FX-trace ``model`` and accept it iff it is a plain chain
    Linear[,ReLU] ... Linear[,log_softmax] whose parameters are exactly the
    Linear weights/biases."""
    import torch.fx as fx

    try:
        gm = fx.symbolic_trace(model)
    except Exception:
        return None
    mods = dict(gm.named_modules())
    layers: List[LinearSpec] = []
    final_lsm = False
    prev = None
    n_placeholder = 0
    for node in gm.graph.nodes:
        if node.op == "placeholder":
            n_placeholder += 1
            if n_placeholder > 1:
                return None
            prev = node
            continue
        if node.op == "output":
            if node.args[0] is not prev:
                return None
            break
        if final_lsm:
            return None                       # nothing may follow the log_softmax
        ins = [a for a in node.args if isinstance(a, fx.Node)]
        if len(ins) != 1 or ins[0] is not prev or len(prev.users) != 1:
            return None
        if node.op == "call_module":
            m = mods[node.target]
            if isinstance(m, nn.Linear):
                layers.append(LinearSpec(m, False))
            elif isinstance(m, nn.ReLU):
                if not layers or layers[-1].relu:
                    return None
                layers[-1].relu = True
            elif isinstance(m, nn.LogSoftmax) and m.dim in (1, -1):
                final_lsm = True
            elif isinstance(m, (nn.Identity, nn.Flatten)):
                pass
            else:
                return None
        elif node.op == "call_function" and node.target in (F.relu, torch.relu):
            if not layers or layers[-1].relu:
                return None
            layers[-1].relu = True
        elif node.op == "call_method" and node.target == "relu":
            if not layers or layers[-1].relu:
                return None
            layers[-1].relu = True
        elif node.op == "call_function" and node.target in (F.log_softmax, torch.log_softmax):
            dim = node.kwargs.get("dim", node.args[1] if len(node.args) > 1 else None)
            if dim not in (1, -1):
                return None
            final_lsm = True
        else:
            return None
        prev = node
    if not layers or layers[-1].relu:
        return None
    covered = set()
    for ls in layers:
        covered.add(id(ls.module.weight))
        if ls.module.bias is not None:
            covered.add(id(ls.module.bias))
    params = [p for p in model.parameters() if p.requires_grad]
    if len(covered) != len(params) or any(id(p) not in covered for p in params):
        return None
    if len(list(model.buffers())) > 0:
        return None
    for a, b in zip(layers[:-1], layers[1:]):
        if a.module.out_features != b.module.in_features:
            return None
    return Lowered(layers, final_lsm)


def loss_kind(criterion, lowered: Lowered) -> Optional[str]:
    if isinstance(criterion, nn.CrossEntropyLoss):
        if criterion.weight is not None or criterion.reduction != "mean" or getattr(criterion, "label_smoothing", 0.0):
            return None
        return "ce"          # CE(log_softmax(z)) == CE(z): a trailing log_softmax is absorbed
    if isinstance(criterion, nn.NLLLoss):
        if not lowered.final_log_softmax or criterion.weight is not None or criterion.reduction != "mean":
            return None
        return "ce"
    if isinstance(criterion, nn.MSELoss):
        if criterion.reduction != "mean" or lowered.final_log_softmax:
            return None
        return "mse"
    return None


# ----------------------------------------------------------------------------- plan
class FusedMLP:
    """Device buffers + enqueue functions of one replica."""

    def __init__(self, model: nn.Module, lowered: Lowered, kind: str, spec, world: int, rank: int,
                 device: torch.device, batch: int, val_rows: int, patience: int, max_iters: int, backend_factory=None,
                 ops=None, act_dtype: torch.dtype = torch.bfloat16):
        """``ops`` / ``act_dtype`` (and a ``backend_factory`` whose product sets ``wants_segments``) exist for the CPU
        emulation of the plan (ops/reference.py, runtime/emulation.py); production: native extension, bf16."""
        from sparktorch_b200.parallel.p2p import FusedUpdate

        self.C = ops if ops is not None else native.load(required=True)
        self.act_dtype = act_dtype
        self.device, self.world, self.rank = device, world, rank
        self.kind, self.lowered = kind, lowered
        self.B, self.nv = batch, val_rows
        L = len(lowered.layers)
        self.L = L
        self.d = [lowered.layers[0].module.in_features] + [ls.module.out_features for ls in lowered.layers]
        d = self.d
        # ---- arenas (symmetric) -------------------------------------------------
        numel = FlatArena.measure(model)
        w_off, off = [], 0
        for l in range(1, L + 1):
            w_off.append(off)
            off += _pad(d[l] * _pad(d[l - 1]), 64)
        fuse_loss = d[L] <= 32 and os.environ.get("SPARKTORCH_FUSE_LOSS", "1") == "1"
        if backend_factory is not None:      # hogwild: owner-sharded parameters, K-AG forward, remote-apply push
            self.fu = backend_factory(numel, off)
        else:
            # Default: two-shot K-AR (owner reduces, publishes -> every replica holds bit-identical parameters and takes
            # the same early-stop decision).  SPARKTORCH_EXCHANGE=push opts into the one-shot exchange where the GEMM
            # epilogues mirror their gradient writes into receive slots on every peer: one barrier less, but the
            # mirrored fp32 atomics land in a different order per copy, so replicas agree only to rounding - it is
            # therefore never combined with early stopping (a split stop vote would dead-lock the barrier).
            want = os.environ.get("SPARKTORCH_EXCHANGE", "twoshot")
            push = world > 1 and fuse_loss and want == "push" and patience <= 0
            from sparktorch_b200.runtime.mega_mlp import mega_extra_elems

            extra = mega_extra_elems(d, [ls.module.bias is not None for ls in lowered.layers], batch)
            self.fu = FusedUpdate(numel, spec, world, rank, device, patience, max_iters, wbf16_elems=off,
                                  exchange="push" if push else "twoshot", grad_extra=extra)
        self.arena = FlatArena(model, device, grad_buffer=self.fu.grad, param_buffer=self.fu.param)
        pos = {id(p): o for p, o in zip(self.arena.params, self.arena.offsets)}
        rows, self.Wb, self.bias, self.gW, self.gb = [], [], [], [], []
        self.peer_kw = []                     # per layer: K-AG arguments of the forward GEMM (empty dict = local weights)
        seg_of = {}
        for l, ls in enumerate(lowered.layers, start=1):
            ldw = _pad(d[l - 1])
            wo = pos[id(ls.module.weight)]
            seg_of[wo] = [wo, d[l], d[l - 1], w_off[l - 1], ldw, -1, 0, 1]
            self.Wb.append(self.fu.wbf16[w_off[l - 1]: w_off[l - 1] + d[l] * ldw].view(d[l], ldw)[:, : d[l - 1]])
            self.gW.append(self.fu.grad[wo: wo + d[l] * d[l - 1]].view(d[l], d[l - 1]))
            bo = -1
            if ls.module.bias is not None:
                bo = pos[id(ls.module.bias)]
                seg_of[bo] = [bo, 1, d[l], -1, 0, -1, 0, 1]
                self.bias.append(self.fu.param[bo: bo + d[l]])
                self.gb.append(self.fu.grad[bo: bo + d[l]])
            else:
                self.bias.append(None)
                self.gb.append(None)
            peer = getattr(self.fu, "peer_table", None)
            if peer is not None:
                self.peer_kw.append(dict(peer_table=peer, peer_shard_elems=self.fu.shard_elems, peer_w_off=wo,
                                         peer_bias_off=bo, peer_fill_local=True))
            else:
                self.peer_kw.append({})
        for o in sorted(seg_of):
            rows.append(seg_of[o])
        self.segment_rows, self.arena_pos = rows, pos          # (K-STEP task builder: runtime/mega_mlp.py)
        if backend_factory is None or getattr(self.fu, "wants_segments", False):
            self.fu.build(rows, zero_grads=True)
        self.loss_slot = self.fu.grad[numel: numel + 1]          # tail[0]: rides in the all-reduce
        self.train_loss_slot = self.fu.grad[numel + 2: numel + 3]  # tail[2]: local record when validating
        # ---- activations ----------------------------------------------------------
        self.train = self._alloc_acts(batch, backward=True)
        self.val = self._alloc_acts(val_rows, backward=False) if val_rows > 0 else None
        self.n_launch = 0
        self.fuse_loss = fuse_loss
        self.window = None            # (n_windows, window_rows): batch = window of the resident partition
        self.side = (torch.cuda.Stream(device=device)
                     if device.type == "cuda" and os.environ.get("SPARKTORCH_SIDE_STREAM", "1") == "1" else None)

    def _alloc_acts(self, rows: int, backward: bool):
        d, L, dev, adt = self.d, self.L, self.device, self.act_dtype
        acts = [torch.zeros(rows, _pad(d[0]), dtype=adt, device=dev)[:, : d[0]]]
        for l in range(1, L):
            acts.append(torch.zeros(rows, _pad(d[l]), dtype=adt, device=dev)[:, : d[l]])
        logits = torch.zeros(rows, _pad(d[L], 16), dtype=torch.float32, device=dev)[:, : d[L]]
        dz = None
        if backward:
            dz = [None] + [torch.zeros(rows, _pad(d[l]), dtype=adt, device=dev)[:, : d[l]] for l in range(1, L + 1)]
        if self.kind == "ce":
            tgt = torch.zeros(rows, dtype=torch.int32, device=dev)
        else:
            tgt = torch.zeros(rows, d[L], dtype=torch.float32, device=dev)
        return {"acts": acts, "logits": logits, "dz": dz, "tgt": tgt, "rows": rows}

    # ---- enqueue helpers (everything goes to the current stream) -----------------
    def _gemm(self, A, a_mn, Bm, b_mn, M, N, K, **kw):
        self.C.gemm_bf16(A, a_mn, Bm, b_mn, M, N, K, **kw)
        self.n_launch += 1

    def _push(self, mask: int) -> dict:
        f = getattr(self.fu, "push_kw", None)
        return f(mask) if f is not None else {}

    def _win(self, flags: int) -> dict:
        if self.window is None or flags == 0:
            return {}
        return dict(win_state=self.fu.state, window_rows=self.window[1], n_windows=self.window[0], win_flags=flags)

    def forward(self, buf, slot=None, train: bool = False, windowed: bool = False) -> None:
        """Forward pass; with ``slot`` the loss (and, when training, dZ_L / db_L) is produced by the
        epilogue of the last GEMM if it fits (out_features <= 32), else by the stand-alone loss kernel."""
        acts, d, L, rows = buf["acts"], self.d, self.L, buf["rows"]
        for l in range(1, L + 1):
            last = l == L
            flags = (1 if (windowed and l == 1) else 0)
            kw = dict(bias=self.bias[l - 1], relu=self.lowered.layers[l - 1].relu)
            if self.peer_kw[l - 1] and buf is self.train:
                # K-AG: this GEMM gathers W_l (and b_l) from the parameter owners over NVLink, casts to bf16 in
                # flight and leaves the gathered copy in the local bf16 arena for the dgrad GEMM
                kw["bias"] = None
                kw.update(self.peer_kw[l - 1])
            if not last:
                kw["out_bf16"] = acts[l]
            elif slot is not None and self.fuse_loss:
                flags |= 4 if windowed else 0
                kw.update(loss_kind=1 if self.kind == "ce" else 2, loss_slot=slot,
                          loss_scale=(1.0 / rows) if self.kind == "ce" else 1.0 / (rows * d[L]))
                kw["labels" if self.kind == "ce" else "target"] = buf["tgt"]
                if train:
                    kw.update(out_bf16=buf["dz"][L], colsum=self.gb[L - 1])
                kw.update(self._push(4 | (2 if train else 0)))      # loss (+ last bias gradient) go to the peers too
            else:
                kw["out_f32"] = buf["logits"]
            kw.update(self._win(flags))
            self._gemm(acts[l - 1], False, self.Wb[l - 1], False, rows, d[l], d[l - 1], **kw)
        if slot is not None and not self.fuse_loss:
            self.loss(buf, slot, train)

    def loss(self, buf, slot, train: bool) -> None:
        d, L = self.d, self.L
        dz = buf["dz"][L] if train else None
        db = self.gb[L - 1] if train else None
        if self.kind == "ce":
            self.C.softmax_ce(buf["logits"], buf["tgt"], d[L], slot, dz, None, db, _pad(d[L]))
        else:
            self.C.mse_loss(buf["logits"], buf["tgt"], d[L], slot, dz, None, db, _pad(d[L]))
        self.n_launch += 1

    def _split_k(self, M: int, N: int, K: int) -> int:
        tiles = ((M + 127) // 128) * ((N + 63) // 64)
        kb = (K + 63) // 64
        if kb < 8 or tiles >= 148:
            return 1
        return max(1, min(kb // 4, 148 // tiles))

    def backward(self, buf, windowed: bool = False) -> None:
        """dgrad chain on the current stream; the weight-gradient GEMMs only feed the update kernel, so
        they run on a side stream (forked/joined with events, captured into the same graph)."""
        acts, dz, d, L, rows = buf["acts"], buf["dz"], self.d, self.L, buf["rows"]
        side = self.side if (self.side is not None and L > 1) else None
        main = torch.cuda.current_stream(self.device) if side is not None else None
        for l in range(L, 0, -1):
            # wgrad: dW_l[d_l, d_{l-1}] = dZ_l^T A_{l-1}  (reduction over the batch; both operands MN-major)
            sk = self._split_k(d[l], d[l - 1], rows)
            wkw = dict(out_f32=self.gW[l - 1], atomic_f32=sk > 1, split_k=sk)
            wkw.update(self._win(2 if (windowed and l == 1) else 0))
            wkw.update(self._push(1))                                  # GEMM -> all-gather: dW tiles land on every peer
            if side is not None and l > 1:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self._gemm(dz[l], True, acts[l - 1], True, d[l], d[l - 1], rows, **wkw)
            else:
                self._gemm(dz[l], True, acts[l - 1], True, d[l], d[l - 1], rows, **wkw)
            if l > 1:
                # dgrad: dZ_{l-1} = (dZ_l W_l) * (A_{l-1} > 0); db_{l-1} = column sums
                relu_prev = self.lowered.layers[l - 2].relu
                self._gemm(dz[l], False, self.Wb[l - 1], True, rows, d[l - 1], d[l],
                           mask_src=acts[l - 1] if relu_prev else None, out_bf16=dz[l - 1], colsum=self.gb[l - 2],
                           **self._push(2))
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)

    def update(self) -> None:
        self.fu.updater.step()
        self.n_launch += 1

    # ---- inference --------------------------------------------------------------------
    def logits_of(self, x_bf16: torch.Tensor) -> torch.Tensor:
        rows = x_bf16.shape[0]
        buf = self._alloc_acts(rows, backward=False)
        buf["acts"][0].copy_(x_bf16)
        self.forward(buf)
        return buf["logits"]


# ----------------------------------------------------------------------------- training driver
def _to_bf16_padded(x: torch.Tensor, device: torch.device, pin: bool = False) -> torch.Tensor:
    """float [n, d] -> bf16 [n, pad8(d)] (zero padded); returns the [:, :d] view."""
    n, d = x.shape
    ld = _pad(d)
    if pin:
        out = torch.zeros(n, ld, dtype=torch.bfloat16).pin_memory()
        out[:, :d].copy_(x)
        return out[:, :d]
    out = torch.zeros(n, ld, dtype=torch.bfloat16, device=device)
    step = max(1, (64 << 20) // max(1, d * 4))
    for s in range(0, n, step):
        out[s: s + step, :d].copy_(x[s: s + step].to(device, non_blocking=True))
    return out[:, :d]


def try_fused_fit(rank, world, model, criterion, opt_cls, opt_params, X, Y, cfg, device, rng):
    from sparktorch_b200.runtime.engine import FitResult, WindowTimer
    from sparktorch_b200.util import split_validation

    if cfg.precision not in ("auto", "bf16") or not native.available():
        return None
    if getattr(cfg, "comm", "auto") not in ("auto", "p2p"):
        return None                     # an explicit collective (NCCL/gloo oracle) exchange was requested
    want_ckpt = bool(cfg.checkpoint_dir and (cfg.checkpoint_every > 0 or cfg.resume))
    if want_ckpt and os.environ.get("SPARKTORCH_MEGA", "1") != "1":
        # mid-training checkpoints / resume on the fused path are a feature of the K-STEP loop (below); without it the
        # eager engine (same update kernel, its optimizer shard and step state are exported) takes over
        return None
    spec = parse_optimizer(opt_cls, opt_params)
    if spec.kind == "other" or world > 8:
        return None
    lowered = lower_mlp(model)
    is_cnn = False
    if lowered is None:
        from sparktorch_b200.runtime.fused_cnn import lower_cnn

        lowered = lower_cnn(model, X.shape[1])
        is_cnn = lowered is not None
    is_resnet = False
    if lowered is None:
        from sparktorch_b200.runtime.fused_resnet import lower_resnet

        lowered = lower_resnet(model, X.shape[1])
        is_resnet = lowered is not None
    if lowered is None:
        return None
    kind = loss_kind(criterion, lowered)
    if kind is None or (is_resnet and kind != "ce"):
        return None
    if is_resnet:
        d0, dL = X.shape[1], lowered.fc.out_features
    elif is_cnn:
        d0, dL = X.shape[1], lowered.linears[-1].module.out_features
    else:
        d0, dL = lowered.layers[0].module.in_features, lowered.layers[-1].module.out_features
    if X.shape[1] != d0:
        return None
    if kind == "mse":
        if Y is None:
            Y = X                                   # unsupervised: target = input
        if Y.shape[1] != dL:
            return None
    elif Y is None:
        return None
    t_start = time.time()
    torch.cuda.set_device(device)
    C = native.load(required=True)

    # ---- data --------------------------------------------------------------------------
    xt = torch.from_numpy(np.ascontiguousarray(X))
    yt = torch.from_numpy(np.ascontiguousarray(Y))
    x_val = y_val = None
    if cfg.validation_pct > 0:
        tr, va = split_validation(X.shape[0], cfg.validation_pct, rng)
        if len(va) > 0:
            x_val, y_val = xt[torch.from_numpy(va)], yt[torch.from_numpy(va)]
            xt, yt = xt[torch.from_numpy(tr)], yt[torch.from_numpy(tr)]
    n = xt.shape[0]
    use_mb = 0 < cfg.mini_batch < n
    Bt = cfg.mini_batch if use_mb else n
    streaming = bool(cfg.streaming and use_mb)
    ytd = (yt.reshape(-1).to(torch.int32) if kind == "ce" else yt.to(torch.float32).reshape(n, -1))
    nv = 0 if x_val is None else x_val.shape[0]

    model = model.to(device)
    if is_resnet:
        from sparktorch_b200.runtime.fused_resnet import FusedResNet

        plan = FusedResNet(model, lowered, kind, spec, world, rank, device, Bt, nv, cfg.early_stop_patience, cfg.iters + 1)
    elif is_cnn:
        from sparktorch_b200.runtime.fused_cnn import FusedCNN

        plan = FusedCNN(model, lowered, kind, spec, world, rank, device, Bt, nv, cfg.early_stop_patience, cfg.iters + 1,
                        seed=0 if cfg.seed is None else int(cfg.seed))
    else:
        plan = FusedMLP(model, lowered, kind, spec, world, rank, device, Bt, nv, cfg.early_stop_patience, cfg.iters + 1)
    if world > 1:
        import torch.distributed as dist

        dist.broadcast(plan.arena.flat_p, src=0)      # rank 0's weights win (quirk Q7)
    stream = torch.cuda.Stream(device=device)
    # staging slots for streamed batches: with the K-STEP kernel a launch runs `stream_spl` steps, the copy stream stays a few
    # launches ahead (3 launches worth of slots)
    stream_spl = max(1, int(os.environ.get("SPARKTORCH_STREAM_STEPS", "4")))
    mlp_candidate = lower_mlp(model) is not None and os.environ.get("SPARKTORCH_MEGA", "1") == "1"
    ring = 3 * stream_spl if (mlp_candidate and cfg.streaming and cfg.early_stop_patience <= 0 and not cfg.verbose) else 3
    runner = C.GraphRunner()
    seed = int(rng.integers(0, 2**62)) if cfg.seed is None else int(cfg.seed) * 1000003 + rank
    with torch.cuda.stream(stream):
        plan.fu.updater.publish()
        sampling = getattr(cfg, "sampling", "random")
        if streaming:
            x_host = _to_bf16_padded(xt, device, pin=True)
            y_host = ytd.pin_memory()
            xs = torch.zeros(ring * Bt, _pad(d0), dtype=torch.bfloat16, device=device)[:, :d0]
            ys = torch.zeros((ring * Bt,) + tuple(ytd.shape[1:]), dtype=ytd.dtype, device=device)
            src, tgt_src, n_src = xs, ys, ring * Bt
            windows = ring                       # the batch of iteration i sits in staging slot i % ring
        else:
            src = _to_bf16_padded(xt, device)
            tgt_src = ytd.to(device)
            n_src = n
            windows = max(1, n // Bt)
        # 'sweep': the batch of iteration i is window i % windows of the (shuffled at ingest) resident
        # partition, addressed inside the GEMMs from the device-side iteration counter - no gather kernel.
        # 'random': a fresh sample without replacement every iteration (the reference's semantics).
        windowed = use_mb and (streaming or sampling == "sweep") and plan.fuse_loss
        gathered = use_mb and not windowed
        gmode = 2 if streaming else (1 if sampling != "sweep" else 0)
        if nv:
            plan.val["acts"][0].copy_(_to_bf16_padded(x_val, device))
            plan.val["tgt"].copy_((y_val.reshape(-1).to(torch.int32) if kind == "ce" else y_val.to(torch.float32).reshape(nv, -1)).to(device))
        if not gathered:                             # full batch or windowed: GEMMs read the partition directly
            plan.train["acts"][0] = src
            plan.train["tgt"] = tgt_src
        if windowed:
            plan.window = (windows, Bt)

        # ---- K-STEP: the whole step (and many steps) as ONE persistent kernel launch (csrc/mlp_step.cu)
        mega = None
        if (not is_cnn and not is_resnet and nv == 0 and plan.fuse_loss and getattr(plan.fu, "exchange", "") == "twoshot"
                and os.environ.get("SPARKTORCH_MEGA", "1") == "1"):
            from sparktorch_b200.parallel.p2p import opt_state_elems
            from sparktorch_b200.runtime.mega_mlp import MegaPlan

            fu = plan.fu
            per = opt_state_elems(fu.numel, world)
            owned = (min(fu.numel, per * rank), min(fu.numel, per * (rank + 1)))
            blk = fu.block
            peers = {k: blk.peer_ptrs(k) for k in ("flags", "grad", "param", "wbf16")}
            if world > 1 and fu.use_multicast:
                peers["mc"] = {k: blk.multicast_ptr(k) for k in ("grad", "param", "wbf16", "flags")}
            flag_view = blk.view("flags", torch.int32, C.flag_words())
            if windowed:
                mx, mt_, mwin = src, tgt_src, (windows, Bt)
            elif gathered:
                mx, mt_, mwin = plan.train["acts"][0], plan.train["tgt"], None
            else:
                mx, mt_, mwin = src, tgt_src, None
            mp = MegaPlan(plan, spec, cfg.early_stop_patience, per * rank, owned, mwin, mx, mt_, flag_view, peers=peers,
                          grid=int(os.environ.get("SPARKTORCH_MEGA_GRID", "148")))
            if mp.ok and not (want_ckpt and (streaming or gathered)):
                mega = mp
                mega_step = mp.materialize()
                if os.environ.get("SPARKTORCH_MEGA_PROF"):      # per-task globaltimer stamps of the last step of a launch
                    mega_prof = torch.zeros(mp.grid_used * 24 * 8, dtype=torch.int64, device=device)
                    mega_step.set_prof(mega_prof)
                    mega_prof_kb = torch.zeros(mp.grid_used * 24 * 64, dtype=torch.int64, device=device)
                    mega_step.set_prof_kb(mega_prof_kb)

        if want_ckpt and mega is None:
            return None       # (CNN / ResNet plans, validation splits, host-side sampling: the eager engine checkpoints those)

        def enqueue_step():
            plan.n_launch = 0
            if gathered:
                C.gather_batch(src, n_src, tgt_src, 0 if kind == "ce" else ytd.shape[1], gmode, ring, seed, Bt, d0,
                               plan.train["acts"][0], None, plan.train["tgt"], plan.fu.state)
                plan.n_launch += 1
            plan.forward(plan.train, slot=plan.train_loss_slot if nv else plan.loss_slot, train=True, windowed=windowed)
            plan.backward(plan.train, windowed=windowed)
            if nv:
                plan.forward(plan.val, slot=plan.loss_slot, train=False)
            plan.update()

        if mega is not None:
            # nothing to warm up or rewind: the kernel is launched directly (or, when every step needs host-side work -
            # H2D streaming, a gather kernel - as a 1-step graph that takes its step numbers from the device state)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            use_pdl = False
            mega_graph = streaming or gathered
            mega_step.run(0, 0, 0, 0)                    # loads the module / sets the shared-memory attribute (no launch)
            if mega_graph:
                if gathered:                             # module load of the gather kernel must not happen inside a capture
                    C.gather_batch(src, n_src, tgt_src, 0 if kind == "ce" else ytd.shape[1], gmode, ring, seed, Bt, d0,
                                   plan.train["acts"][0], None, plan.train["tgt"], plan.fu.state)
                    stream.synchronize()
                runner.begin_capture()
                if gathered:
                    C.gather_batch(src, n_src, tgt_src, 0 if kind == "ce" else ytd.shape[1], gmode, ring, seed, Bt, d0,
                                   plan.train["acts"][0], None, plan.train["tgt"], plan.fu.state)
                mega_step.run(-1, 0, 0, 1)
                runner.end_capture()
                if streaming and ring == 3 * stream_spl and stream_spl > 1:
                    # main graph: stream_spl steps per launch (the one-step graph above serves ragged tails)
                    runner.keep_as_tail_and_expect(stream_spl)
                    runner.begin_capture()
                    mega_step.run(-1, 0, 0, stream_spl)
                    runner.end_capture()
            launches = 1 + (1 if gathered else 0)
        if mega is None:
            # warm-up launch outside capture (module load, smem attributes), then rewind the step state
            p0 = plan.fu.param.clone()
            if hasattr(plan, "save_extra"):
                plan.save_extra()                         # e.g. batch-norm running statistics
            if streaming:
                xs[:Bt].copy_(x_host[:Bt].to(device))
                ys[:Bt].copy_(y_host[:Bt].to(device))
            enqueue_step()
            stream.synchronize()
            plan.fu.param.copy_(p0)
            if hasattr(plan, "restore_extra"):
                plan.restore_extra()
            for t in (plan.fu.m, plan.fu.v):
                if t is not None:
                    t.zero_()
            plan.fu.state[2:].zero_()                     # keep epoch/blocks_done, reset step/iter/early-stop record
            plan.fu.loss_hist.zero_()
            plan.fu.mailbox.zero_()
            plan.fu.grad.zero_()
            plan.fu.updater.publish()
            stream.synchronize()
            if world > 1:
                dist.barrier()
            # Capture the step.  Kernels are launched with programmatic dependent launch so the prologue of
            # kernel n+1 (barrier init, TMEM alloc, descriptor prefetch) overlaps the tail of kernel n; if the
            # driver refuses such a capture we fall back to plain stream-ordered edges.
            use_pdl = os.environ.get("SPARKTORCH_PDL", "1") == "1"
            for attempt in range(2):
                C.set_pdl(use_pdl)
                try:
                    runner.begin_capture()
                    enqueue_step()
                    runner.end_capture()
                    break
                except RuntimeError:
                    C.set_pdl(False)
                    if not use_pdl or attempt == 1:
                        raise
                    try:
                        runner.end_capture()
                    except RuntimeError:
                        pass
                    runner = C.GraphRunner()
                    use_pdl = False
            C.set_pdl(False)
            launches = plan.n_launch
        if streaming:
            runner.setup_streaming(x_host.data_ptr(), y_host.data_ptr(), (n // Bt) * Bt, x_host.stride(0) * 2,
                                   y_host.element_size() * (y_host[0].numel() if y_host.dim() > 1 else 1),
                                   xs.data_ptr(), ys.data_ptr(), Bt, ring)

        # ---- the loop: native, no Python per iteration ------------------------------------
        window = WindowTimer(cfg.time_window, device, world)
        mail_ptr = plan.fu.mailbox.data_ptr()
        check = 1 if cfg.verbose else (4 if cfg.early_stop_patience > 0 else 0)
        steps_per_launch = int(os.environ.get("SPARKTORCH_MEGA_STEPS", "8" if cfg.early_stop_patience > 0 else "64"))
        t_loop = time.time()
        done = 0
        start_iter = 0
        if want_ckpt:
            from sparktorch_b200.utils import checkpoint as ckpt

            if cfg.resume:
                # the ranks agree on the newest iteration every one of them holds (utils/checkpoint.py); the step kernel takes
                # its iteration / optimizer-step numbers as launch parameters, its dependency counters start fresh
                gens = {it: pl for it, pl in ckpt.load_generations(cfg.checkpoint_dir, rank).items()
                        if pl.get("numel") == plan.fu.numel and pl.get("engine") == "kstep" and pl.get("world") == world}
                common = ckpt.agree_on_iteration(sorted(gens), world, device)
                if common > 0:
                    sv = gens[common]
                    fu = plan.fu
                    fu.param.copy_(sv["param"].to(device))
                    for dst, key in ((fu.m, "m"), (fu.v, "v")):
                        if dst is not None and sv.get(key) is not None:
                            dst.copy_(sv[key].to(device))
                    fu.state.copy_(sv["state"].to(device))
                    fu.state[11] = 0                      # StepState.pad[0]: steps since the counters were zeroed
                    nh = min(fu.loss_hist.numel(), sv["loss_hist"].numel())
                    fu.loss_hist[:nh].copy_(sv["loss_hist"][:nh].to(device))
                    fu.mailbox[1] = int(common)
                    fu.updater.publish()                  # bf16 operand copies of the restored weights
                    stream.synchronize()
                    start_iter = int(common)
                if world > 1:
                    dist.barrier()
        segments = [(start_iter, cfg.iters)]
        if cfg.time_window is not None:
            w0, wk = cfg.time_window
            segments = [(0, w0), (w0, w0 + wk), (w0 + wk, cfg.iters)]
        if want_ckpt and cfg.checkpoint_every > 0:          # cut every segment at the checkpoint boundaries
            cut = []
            for lo, hi in segments:
                while lo < hi:
                    nxt = min(hi, (lo // cfg.checkpoint_every + 1) * cfg.checkpoint_every)
                    cut.append((lo, nxt))
                    lo = nxt
            segments = cut
        printed = 0
        from sparktorch_b200.utils.watchdog import ProgressWatchdog

        mailbox = plan.fu.mailbox
        wd = ProgressWatchdog(lambda: int(mailbox[1]), label=f"rank {rank} fused loop")
        wd.__enter__()
        try:
            for seg_i, (lo, hi) in enumerate(segments):
                if hi <= lo:
                    continue
                timed = cfg.time_window is not None and seg_i == 1
                if timed:
                    window.before_step(lo)
                if cfg.verbose:
                    for i in range(lo, hi):
                        got = (mega_step.run_many(i, i, i - start_iter, 1, 1, mail_ptr) if (mega is not None and not mega_graph)
                               else runner.run(i, 1, mail_ptr, 1))
                        stream.synchronize()
                        done += got
                        print(f"Partition: {rank}. Iteration: {i}. Distributed Loss: {plan.fu.last_loss()}")
                        printed += 1
                        if plan.fu.stopped():
                            break
                elif mega is not None and not mega_graph:
                    # steps_per_launch whole steps per kernel launch; the kernel itself idles through the rest of a launch
                    # once the stop is latched, the host looks at the mailbox between launches
                    done += mega_step.run_many(lo, lo, lo - start_iter, hi - lo, steps_per_launch, mail_ptr if check else 0)
                else:
                    done += runner.run(lo, hi - lo, mail_ptr, check)
                if timed:
                    window.after_step(hi - 1)
                if plan.fu.stopped():
                    break
                if want_ckpt and cfg.checkpoint_every > 0 and hi % cfg.checkpoint_every == 0:
                    stream.synchronize()
                    fu = plan.fu
                    ckpt.save(cfg.checkpoint_dir, rank, {
                        "engine": "kstep", "world": world, "numel": fu.numel, "iters_done": hi, "param": fu.param.cpu(),
                        "m": None if fu.m is None else fu.m.cpu(), "v": None if fu.v is None else fu.v.cpu(),
                        "state": fu.state.cpu(), "loss_hist": fu.loss_hist.cpu()})
                    _inject = os.environ.get("SPARKTORCH_FAULT_INJECT")
                    if _inject:
                        from sparktorch_b200.runtime.engine import _maybe_inject_fault

                        _maybe_inject_fault(cfg, rank, hi)
            stream.synchronize()
        finally:
            wd.__exit__()
    torch.cuda.synchronize(device)
    if mega is not None and int(mega.d_error[0]) != 0:
        raise RuntimeError("K-STEP kernel gave up waiting (code, cta, task, epoch) = %s" % mega.d_error.cpu().tolist())
    if mega is not None and os.environ.get("SPARKTORCH_MEGA_PROF"):
        import json

        stamps = mega_prof.cpu().view(mega.grid_used, 24, 8).tolist()
        kbs = mega_prof_kb.cpu().view(mega.grid_used, 24, 16, 4).tolist()
        recs = []
        for ti, td in enumerate(mega.task_dicts):
            c, j = mega.cta_of_task[ti]
            recs.append({"group": td["group"], "cta": c, "slot": j, "m0": td["m0"], "n0": td["n0"], "t": stamps[c][j],
                         "kb": kbs[c][j][: max(1, min(16, td["num_kb"]))]})
        with open(os.environ["SPARKTORCH_MEGA_PROF"] + ("" if rank == 0 else f".rank{rank}"), "w") as fh:
            json.dump({"world": world, "tasks": recs, "info": mega.describe()}, fh)
    t_end = time.time()
    n_done = plan.fu.iters_done()
    if hasattr(plan, "finish"):
        plan.finish(n_done)
    stopped = plan.fu.stopped()
    losses = plan.fu.loss_hist[:n_done].cpu().tolist()
    history = {
        "engine": "fused_resnet" if is_resnet else ("fused_cnn" if is_cnn else "fused_mlp"),
        "step_kernel": (dict(mega.describe(), steps_per_launch=((stream_spl if (streaming and ring == 3 * stream_spl) else 1)
                                                                if mega_graph else steps_per_launch)) if mega is not None else None),
        "comm": ("P2PPush" if getattr(plan.fu, "exchange", "") == "push" else "P2PFused" + ("+NVLS" if plan.fu.use_multicast else "")),
        "precision": "bf16",
        "world_size": world, "train_loss": losses, "stopped_at": (n_done - 1) if stopped else None, "iters_run": n_done,
        "resumed_from": start_iter,
        "samples_per_step": Bt * world, "loop_seconds": t_end - t_loop, "total_seconds": t_end - t_start,
        "window_ms": window.result_ms(), "gpu_launches_per_step": launches, "graph_nodes": runner.num_nodes() if (mega is None or mega_graph) else 0,
        "streaming": streaming, "sampling": ("stream" if streaming else sampling) if use_mb else "full-batch",
        "fused_loss": plan.fuse_loss, "pdl": use_pdl,
        "h2d_bytes_per_step": (Bt * (_pad(d0) * 2 + (4 if kind == "ce" else 4 * dL))) if streaming else 0,
        "d2h_bytes_per_step": 12,          # loss / iteration count / stop flag stored to the pinned mailbox every step
    }
    state = plan.arena.detach_state_dict() if rank == 0 else None
    return FitResult(state, history)


# ----------------------------------------------------------------------------- inference
def try_fused_predictor(model: nn.Module, device: torch.device, precision: str = "auto") -> Optional[Callable]:
    if precision not in ("auto", "bf16") or not native.available():
        return None
    lowered = lower_mlp(model)
    if lowered is None:
        return _ConvPredictor(model, device) if any(isinstance(m, nn.Conv2d) for m in model.modules()) else None
    C = native.load(required=True)
    d = [lowered.layers[0].module.in_features] + [ls.module.out_features for ls in lowered.layers]
    Wb, bias = [], []
    for l, ls in enumerate(lowered.layers, start=1):
        w = torch.zeros(d[l], _pad(d[l - 1]), dtype=torch.bfloat16, device=device)
        w[:, : d[l - 1]].copy_(ls.module.weight.detach())
        Wb.append(w[:, : d[l - 1]])
        bias.append(None if ls.module.bias is None else ls.module.bias.detach().float().contiguous())
    L = len(lowered.layers)

    def run(x: torch.Tensor) -> torch.Tensor:
        rows = x.shape[0]
        a = torch.zeros(rows, _pad(d[0]), dtype=torch.bfloat16, device=device)[:, : d[0]]
        a.copy_(x)
        for l in range(1, L + 1):
            last = l == L
            if last:
                out = torch.zeros(rows, _pad(d[l], 16), dtype=torch.float32, device=device)[:, : d[l]]
                C.gemm_bf16(a, False, Wb[l - 1], False, rows, d[l], d[l - 1], 1.0, bias[l - 1], False, None, None, None, out,
                            None, False, None, 0, 1)
                return torch.log_softmax(out, dim=1) if lowered.final_log_softmax else out
            nxt = torch.zeros(rows, _pad(d[l]), dtype=torch.bfloat16, device=device)[:, : d[l]]
            C.gemm_bf16(a, False, Wb[l - 1], False, rows, d[l], d[l - 1], 1.0, bias[l - 1], lowered.layers[l - 1].relu, None,
                        nxt, None, None, None, False, None, 0, 1)
            a = nxt
        return a

    return run


class _ConvPredictor:
    """Inference through the fused CNN / ResNet plans (forward kernels only).  The input width is only known at the
    first batch, so lowering happens then; nets that do not lower keep running through ``model(x)``."""

    def __init__(self, model: nn.Module, device: torch.device):
        self.model, self.device = model, device
        self.plan = None
        self.failed = False
        self.engine = None

    def _build(self, in_features: int) -> None:
        from sparktorch_b200.runtime.fused_cnn import FusedCNN, lower_cnn
        from sparktorch_b200.runtime.fused_resnet import FusedResNet, lower_resnet

        spec = parse_optimizer(torch.optim.SGD, {"lr": 0.0})
        low = lower_cnn(self.model, in_features)
        if low is not None:
            per_row = sum(g_["OH"] * g_["OW"] * (g_["K"] + g_["Cout"]) for g_ in _cnn_geo(low)) * 2
            self.chunk = int(max(64, min(8192, (1 << 30) // max(1, per_row))))
            kind = "ce" if low.linears[-1].module.out_features > 1 else "mse"
            self.plan = FusedCNN(self.model, low, kind, spec, 1, 0, self.device, self.chunk, 0, 0, 1, inference=True)
            self.engine = "fused_cnn"
        else:
            low = lower_resnet(self.model, in_features)
            if low is None:
                self.failed = True
                return
            self.chunk = 64
            self.plan = FusedResNet(self.model, low, "ce", spec, 1, 0, self.device, self.chunk, 0, 0, 1, inference=True)
            self.engine = "fused_resnet"
        self.low = low
        self.plan.fu.updater.publish()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.plan is None and not self.failed:
            try:
                self._build(int(x.shape[1]))
            except Exception:
                self.failed = True
        if self.failed:
            return self.model(x)
        buf, outs = self.plan.train, []
        for s in range(0, x.shape[0], self.chunk):
            xb = x[s: s + self.chunk]
            buf["acts"][0][: xb.shape[0]].copy_(xb)
            self.plan.forward(buf, slot=None, train=False)
            outs.append(buf["logits"][: xb.shape[0]].clone())
        out = torch.cat(outs, 0) if len(outs) > 1 else outs[0]
        return torch.log_softmax(out, dim=1) if getattr(self.low, "final_log_softmax", False) else out


def _cnn_geo(low):
    c, h, w = low.in_shape
    out = []
    for cs in low.convs:
        m = cs.module
        k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
        st = m.stride if isinstance(m.stride, int) else m.stride[0]
        p = m.padding if isinstance(m.padding, int) else m.padding[0]
        oh, ow = (h + 2 * p - k) // st + 1, (w + 2 * p - k) // st + 1
        out.append(dict(OH=oh, OW=ow, K=c * k * k, Cout=m.out_channels))
        c, h, w = m.out_channels, oh // (cs.pool or 1), ow // (cs.pool or 1)
    return out
