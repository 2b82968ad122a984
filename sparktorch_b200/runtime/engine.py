"""Per-rank training engine (the body of a barrier task).

Replaces the reference's Python hot loop (distributed.py:140-203; call stack in
SURVEY.md §3.1).  Differences by design:

* one flat gradient arena → ONE collective (or one fused P2P kernel) per step,
  loss and stop vote piggy-backed in the same payload (no extra all-reduces,
  reference distributed.py:188,194);
* labels get their dtype once, from the criterion (quirk Q3);
* mini-batches are sampled on the training device (quirk Q2/G12);
* validation runs under ``no_grad`` in eval mode (quirk Q6);
* rank 0's initial weights are broadcast (quirk Q7), float buffers averaged at
  the end (quirk Q8), only rank 0 ships a state_dict back (quirk Q14);
* the per-step loss history is returned to the driver (the reference drops it).

Two step implementations share this loop:

* ``EagerStep``  – autograd over the user's module (any ``nn.Module``), CPU or
  GPU; on B200 the Linear / Conv2d / CE / MSE work is substituted by our kernels
  (ops/module_swap.py) and the whole step is replayed as one CUDA graph
  (``GraphedStep``);
* ``FusedMLPStep`` (runtime/fused_mlp.py) – recognised MLP + CE/MSE + Adam/SGD
  runs as a CUDA graph of hand-written sm_100a kernels, no autograd, no Python
  in the loop.
"""
from __future__ import annotations

import dataclasses
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from sparktorch_b200.runtime.flat import FlatArena
from sparktorch_b200.runtime.optim import parse_optimizer
from sparktorch_b200.util import criterion_wants_class_index, split_validation


@dataclasses.dataclass
class TrainConfig:
    iters: int = 10
    mini_batch: int = -1
    validation_pct: float = 0.0
    early_stop_patience: int = -1
    verbose: int = 0
    device: str = "cpu"
    precision: str = "auto"       # auto | fp32 | bf16
    engine: str = "auto"          # auto | eager | fused
    comm: str = "auto"            # auto | collective | p2p
    seed: Optional[int] = None
    streaming: bool = False       # keep the partition in pinned host memory, H2D one batch per step
    time_window: Optional[Tuple[int, int]] = None   # (first_step, n_steps): device-time this window (bench/profiling)
    sampling: str = "random"      # random: fresh sample w/o replacement per step | sweep: walk the shuffled partition
    checkpoint_dir: Optional[str] = None   # every rank writes <dir>/rank<r>.pt
    checkpoint_every: int = 0              # iterations between checkpoints (0: off)
    resume: bool = False                   # continue from the checkpoint in checkpoint_dir if there is one


@dataclasses.dataclass
class FitResult:
    state_dict: Optional[Dict[str, torch.Tensor]]
    history: Dict[str, Any]


def resolve_device(device: str, local_index: int) -> torch.device:
    if device is None or str(device) == "cpu":
        return torch.device("cpu")
    if str(device).startswith("cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("device='cuda' requested but no CUDA device is visible in this task")
        if ":" in str(device):
            return torch.device(device)
        return torch.device("cuda", local_index % torch.cuda.device_count())
    return torch.device(device)


def prepare_targets(Y: Optional[torch.Tensor], criterion) -> Optional[torch.Tensor]:
    if Y is None:
        return None
    if criterion_wants_class_index(criterion):
        if Y.dim() == 1 or Y.shape[1] == 1:
            return Y.reshape(-1).long()
        # one-hot / probability targets (k columns): CrossEntropyLoss takes them as float class probabilities; the reference
        # reaches the same call because its first criterion(...) try succeeds for them (distributed.py:152-157)
        return Y.to(torch.float32)
    return Y


class WindowTimer:
    """Device-times steps [first, first+n): barrier + synchronize on both sides,
    CUDA events on the compute stream, MAX over ranks."""

    def __init__(self, window, device: torch.device, world: int):
        self.window, self.device, self.world = window, device, world
        self.ms = None
        self._t0 = None
        if window is not None and device.type == "cuda":
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def before_step(self, i: int) -> None:
        if self.window is not None and i == self.window[0]:
            self._sync()
            self._t0 = time.perf_counter()
            self.host_t0 = time.time()
            if self.device.type == "cuda":
                self.e0.record()

    def after_step(self, i: int) -> None:
        if self.window is not None and i == self.window[0] + self.window[1] - 1:
            if self.device.type == "cuda":
                self.e1.record()
                torch.cuda.synchronize(self.device)
                ms = self.e0.elapsed_time(self.e1)
            else:
                ms = (time.perf_counter() - self._t0) * 1e3
            wall = (time.perf_counter() - self._t0) * 1e3
            t = torch.tensor([ms, wall], dtype=torch.float64, device=self.device)
            if self.world > 1:
                import torch.distributed as dist

                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.ms = (float(t[0]), float(t[1]))
            self.host_t1 = time.time()
            self._sync()

    def result_ms(self):
        if self.ms is None:
            return None
        return {"device_ms": self.ms[0], "wall_ms": self.ms[1], "host_t0": getattr(self, "host_t0", None),
                "host_t1": getattr(self, "host_t1", None)}


class EagerStep:
    """fwd → loss → bwd on the user's module; gradients land in the arena."""

    name = "eager"

    def __init__(self, model: nn.Module, criterion, arena: FlatArena, autocast_bf16: bool = False):
        self.model, self.criterion, self.arena = model, criterion, arena
        self.autocast_bf16 = autocast_bf16      # precision='bf16' on the eager GPU path: bf16 compute, fp32 master weights

    def run(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.arena.zero_grad()
        if self.autocast_bf16:
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                out = self.model(x)
            loss = self.criterion(out.float(), y)
        else:
            out = self.model(x)
            loss = self.criterion(out, y)
        loss.backward()
        return loss.detach()

    @torch.no_grad()
    def evaluate(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        was_training = self.model.training
        self.model.eval()
        try:
            return self.criterion(self.model(x), y).detach()
        finally:
            self.model.train(was_training)


class GraphedStep:
    """One training step of the eager engine (any nn.Module under autograd, GEMMs on our kernels through
    ops/module_swap.py, the fused K-AR update) captured into a CUDA graph and replayed: no Python, no per-op launch
    overhead in the loop.  Static input tensors are refilled before every replay.  Falls back (``capture() -> False``)
    when the module cannot be captured (host syncs, data-dependent control flow, ...)."""

    def __init__(self, step: "EagerStep", comm, arena: FlatArena, x_all, y_all, x_val, y_val, batch: int, static_input: bool,
                 device: torch.device):
        self.step, self.comm, self.arena, self.device = step, comm, arena, device
        self.x_val, self.y_val = x_val, y_val
        self.static_input = static_input          # full batch: the resident partition IS the input
        if static_input:
            self.sx, self.sy = x_all, y_all
        else:
            self.sx = torch.zeros((batch,) + tuple(x_all.shape[1:]), dtype=x_all.dtype, device=device)
            self.sy = self.sx if y_all is x_all else torch.zeros((batch,) + tuple(y_all.shape[1:]), dtype=y_all.dtype, device=device)
        self.loss = torch.zeros((), dtype=torch.float32, device=device)
        self.val_loss = torch.zeros((), dtype=torch.float32, device=device) if x_val is not None else None
        self.graph = None

    def _body(self) -> None:
        from sparktorch_b200.ops.module_swap import invalidate_weight_caches

        invalidate_weight_caches()                 # the bf16 weight casts must be part of every replay
        loss = self.step.run(self.sx, self.sy)
        self.loss.copy_(loss)
        if self.x_val is not None:
            v = self.step.evaluate(self.x_val, self.y_val)
            self.val_loss.copy_(v)
            self.arena.tail[0] = v
        else:
            self.arena.tail[0] = loss
        self.comm.fu.updater.step()

    def capture(self, world: int) -> bool:
        fu = self.comm.fu
        saved = {"param": fu.param.clone(), "state": fu.state.clone(), "m": None if fu.m is None else fu.m.clone(),
                 "v": None if fu.v is None else fu.v.clone()}
        bufs = [b.clone() for b in self.step.model.buffers()]
        ok = True
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):                 # warm-up on a side stream (autograd / allocator settle), as torch.cuda.graphs asks for
                    self._body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
        except Exception as e:                     # not capturable: run the plain eager loop
            self.error = repr(e)
            ok = False
            try:
                torch.cuda.synchronize(self.device)
            except Exception:
                pass
        # every rank must take the same path (the captured step contains the cross-GPU exchange)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(int(t.item()))
        # rewind what warm-up and capture changed: weights, optimizer state, step state, loss history, statistics buffers
        with torch.no_grad():
            fu.param.copy_(saved["param"])
            if fu.m is not None:
                fu.m.copy_(saved["m"])
            if fu.v is not None:
                fu.v.copy_(saved["v"])
            fu.state[2:].copy_(saved["state"][2:])
            fu.loss_hist.zero_()
            fu.mailbox.zero_()
            fu.grad.zero_()
            for b, b0 in zip(self.step.model.buffers(), bufs):
                b.copy_(b0)
        torch.cuda.synchronize(self.device)
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        return ok

    def replay(self, xb, yb):
        if not self.static_input:
            self.sx.copy_(xb, non_blocking=True)
            if self.sy is not self.sx:
                self.sy.copy_(yb, non_blocking=True)
        self.graph.replay()
        return self.loss, self.val_loss


class DeviceSampler:
    """Mini-batch index sampler living on the training device (no host sync)."""

    def __init__(self, n: int, batch: int, device: torch.device, seed: Optional[int]):
        self.n, self.batch = n, batch
        self.gen = torch.Generator(device=device)
        if seed is not None:
            self.gen.manual_seed(seed)
        else:
            self.gen.seed()
        self.device = device

    def next(self) -> torch.Tensor:
        # without replacement inside a batch, independent across steps (reference
        # semantics: np.random.choice(n, mb, replace=False), distributed.py:146)
        return torch.randperm(self.n, device=self.device, generator=self.gen)[: self.batch]


def _maybe_inject_fault(cfg: "TrainConfig", rank: int, it: int) -> None:
    """Test hook: SPARKTORCH_FAULT_INJECT='<rank>:<iter>' kills that task once (a marker file in the
    checkpoint directory prevents re-triggering after the stage is retried)."""
    import os

    spec = os.environ.get("SPARKTORCH_FAULT_INJECT")
    if not spec:
        return
    r, _, k = spec.partition(":")
    if int(r) != rank or int(k) != it:
        return
    marker = os.path.join(cfg.checkpoint_dir or "/tmp", f".fault_{r}_{k}")
    if os.path.exists(marker):
        return
    os.makedirs(os.path.dirname(marker), exist_ok=True)
    open(marker, "w").close()
    raise RuntimeError(f"injected fault on rank {rank} at iteration {it}")


def build_comm_and_arena(model: nn.Module, opt_cls, opt_params, cfg: TrainConfig, world: int, rank: int,
                         device: torch.device):
    """Pick the gradient-exchange back-end and lay the model out in its arena."""
    from sparktorch_b200.parallel.comm import CollectiveComm, LocalComm

    spec = parse_optimizer(opt_cls, opt_params)
    patience = cfg.early_stop_patience
    if cfg.comm in ("auto", "p2p") and device.type == "cuda" and spec.kind != "other":
        # The P2P constructor contains collectives (symmetric-memory rendezvous, barrier): either EVERY rank goes that way
        # or none does.  Only the locally checkable precondition (native extension present) may veto it, and the veto is
        # agreed on by all ranks BEFORE anyone enters the constructor; a failure inside it is a real error and is raised
        # (a rank silently falling back to NCCL while its peers spin in the flag barrier would dead-lock the job).
        from sparktorch_b200.ops import native

        ok = 1 if native.available() else 0
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = int(t.item())
        if ok:
            from sparktorch_b200.parallel.p2p import P2PFusedComm

            comm = P2PFusedComm(model, spec, patience, world, rank, device, cfg.iters)
            return comm, comm.arena
        if cfg.comm == "p2p":
            raise RuntimeError("comm='p2p' requested but the native extension is unavailable on at least one rank: " + str(native.error()))
    arena = FlatArena(model, device)
    user_opt = None
    if spec.kind == "other":
        user_opt = opt_cls(arena.params, **opt_params) if opt_params else opt_cls(arena.params)
    comm = LocalComm(arena, spec, user_opt, patience) if world == 1 else CollectiveComm(arena, spec, user_opt, patience)
    return comm, arena


def train_partition(
    rank: int,
    world: int,
    model: nn.Module,
    criterion,
    opt_cls,
    opt_params: Optional[Dict],
    X: np.ndarray,
    Y: Optional[np.ndarray],
    cfg: TrainConfig,
    device: torch.device,
) -> FitResult:
    """Train one replica on its partition; returns rank 0's weights + history."""
    t_start = time.time()
    seed = cfg.seed
    rng = np.random.default_rng(None if seed is None else seed + 7919 * rank)

    # ---- fused path (recognised MLP on B200) --------------------------------
    if cfg.engine in ("auto", "fused") and device.type == "cuda":
        try:
            from sparktorch_b200.runtime.fused_mlp import try_fused_fit

            res = try_fused_fit(rank, world, model, criterion, opt_cls, opt_params, X, Y, cfg, device, rng)
            if res is not None:
                return res
            if cfg.engine == "fused":
                raise RuntimeError("engine='fused' requested but the model/criterion/optimizer is not lowerable "
                                   "(or checkpointing/resume was requested, which only the eager engine implements)")
        except ImportError:
            if cfg.engine == "fused":
                raise

    # ---- data ---------------------------------------------------------------
    xt = torch.from_numpy(X)
    yt = torch.from_numpy(Y) if Y is not None else None
    x_val = y_val = None
    if cfg.validation_pct > 0:
        tr, va = split_validation(X.shape[0], cfg.validation_pct, rng)
        if len(va) > 0:
            x_val = xt[torch.from_numpy(va)]
            y_val = yt[torch.from_numpy(va)] if yt is not None else None
            xt = xt[torch.from_numpy(tr)]
            yt = yt[torch.from_numpy(tr)] if yt is not None else None
    stream_host = cfg.streaming and device.type == "cuda" and 0 < cfg.mini_batch < xt.shape[0]
    if stream_host:       # partition stays in pinned host memory; one batch is copied per step
        x_all = xt.pin_memory()
        y_all = prepare_targets(yt, criterion).pin_memory() if yt is not None else x_all
    else:
        x_all = xt.to(device)
        y_all = prepare_targets(yt.to(device), criterion) if yt is not None else x_all
    if x_val is not None:
        x_val = x_val.to(device)
        y_val = prepare_targets(y_val.to(device), criterion) if y_val is not None else x_val

    # ---- model / arena / comm ----------------------------------------------
    model = model.to(device)
    model.train()
    if hasattr(criterion, "to"):
        criterion = criterion.to(device)
    swap_report = None
    if device.type == "cuda" and cfg.engine != "eager_plain":
        # GEMM-shaped work of ANY module goes to our sm_100a kernels (ops/module_swap.py): Linear / Conv2d forward, dgrad,
        # wgrad on the tcgen05 GEMM, fused CE / MSE; the rest of the module runs as it is under autograd
        from sparktorch_b200.ops.module_swap import swap_modules

        model, criterion, swap_report = swap_modules(model, criterion, precision=cfg.precision)
    swapped = swap_report is not None and (swap_report["linear"] + swap_report["conv2d"]) > 0
    comm, arena = build_comm_and_arena(model, opt_cls, opt_params, cfg, world, rank, device)
    comm.broadcast_params()
    # (swapped modules do their own bf16 casts; torch.autocast is only for the un-swapped eager path)
    step = EagerStep(model, criterion, arena, autocast_bf16=(cfg.precision == "bf16" and device.type == "cuda" and not swapped))

    n = x_all.shape[0]
    use_mb = 0 < cfg.mini_batch < n
    sampler = None
    if use_mb and not stream_host:
        sampler = DeviceSampler(n, cfg.mini_batch, device, None if seed is None else seed + rank)
    window = WindowTimer(cfg.time_window, device, world)
    host_loss = torch.zeros(max(cfg.iters, 1), dtype=torch.float32).pin_memory() if stream_host else None
    h2d_bytes = d2h_bytes = 0
    from sparktorch_b200.utils import checkpoint as ckpt
    from sparktorch_b200.utils.tracing import PhaseTimer, trace_path

    tracer = PhaseTimer(device, enabled=trace_path(rank) is not None)
    has_es = cfg.early_stop_patience > 0
    need_loss = bool(cfg.verbose) or has_es
    losses: List[float] = []
    loss_buf = torch.zeros(max(cfg.iters, 1), dtype=torch.float32, device=device)
    start_iter = 0
    resume_sampler_state = None
    if cfg.resume and cfg.checkpoint_dir:
        # Every rank must resume from the SAME iteration with the same weights: a crash between two ranks' saves, a
        # missing / foreign file or a non-shared directory would otherwise let the ranks run different numbers of
        # exchanges (dead-lock) or train diverged replicas.  Checkpoints keep two generations; the ranks agree on the
        # newest generation that all of them hold, else everybody restarts from iteration 0.
        gens = ckpt.load_generations(cfg.checkpoint_dir, rank)
        gens = {it: pl for it, pl in gens.items() if pl.get("numel") == arena.numel and pl.get("comm_kind", type(comm).__name__) == type(comm).__name__}
        common = ckpt.agree_on_iteration(sorted(gens), world, device)
        if common > 0:
            saved = gens[common]
            with torch.no_grad():
                arena.flat_p.copy_(saved["flat_p"].to(device))
            comm.import_state(saved["comm"])
            start_iter = int(saved["iters_done"])
            loss_buf[:start_iter] = torch.tensor(saved["losses"][:start_iter], device=device)
            resume_sampler_state = saved.get("sampler_state")
        # rank 0's weights win again (they are identical across ranks in every complete generation; this also repairs a
        # fall-back to iteration 0 after some ranks had already loaded something)
        comm.broadcast_params()
    if resume_sampler_state is not None and sampler is not None:
        sampler.gen.set_state(resume_sampler_state.to("cpu"))
    stopped_at = None
    t_loop = time.time()

    # peers of a dead rank would spin in the K-AR flag barrier forever: turn that into a loud task failure
    progress = [start_iter]
    stall_guard = None
    if world > 1 and getattr(comm, "fused", False):
        from sparktorch_b200.utils.watchdog import ProgressWatchdog

        stall_guard = ProgressWatchdog(lambda: progress[0], label=f"rank {rank} eager loop", grace_first_s=600.0)
        stall_guard.__enter__()
    graphed = None
    if (device.type == "cuda" and getattr(comm, "fused", False) and cfg.engine != "eager_plain" and not cfg.checkpoint_dir
            and __import__("os").environ.get("SPARKTORCH_EAGER_GRAPH", "1") == "1"):
        bsz = cfg.mini_batch if use_mb else n
        graphed = GraphedStep(step, comm, arena, x_all, y_all, x_val, y_val, bsz, static_input=not use_mb, device=device)
        if not graphed.capture(world):
            graphed = None
    try:
        for i in range(start_iter, cfg.iters):
            progress[0] = i
            _maybe_inject_fault(cfg, rank, i)
            window.before_step(i)
            if stream_host:
                # sequential sweep over the (already shuffled) pinned partition
                lo = (i * cfg.mini_batch) % (n - cfg.mini_batch + 1)
                xb = x_all[lo:lo + cfg.mini_batch].to(device, non_blocking=True)
                yb = xb if y_all is x_all else y_all[lo:lo + cfg.mini_batch].to(device, non_blocking=True)
                h2d_bytes = xb.numel() * xb.element_size() + (0 if yb is xb else yb.numel() * yb.element_size())
            elif sampler is not None:
                idx = sampler.next()
                xb, yb = x_all[idx], (y_all[idx] if y_all is not x_all else None)
                if yb is None:
                    yb = xb
            else:
                xb, yb = x_all, y_all
            if graphed is not None:
                # the whole step (autograd forward/backward through the swapped kernels, validation forward, K-AR) is ONE graph
                loss, val_loss = graphed.replay(xb, yb)
                loss_buf[i].copy_(loss)
                res = comm.after_graph_step(need_loss)
            else:
                with tracer.phase("forward_backward"):
                    loss = step.run(xb, yb)
                val_loss = step.evaluate(x_val, y_val) if x_val is not None else None
                # piggy-back: loss used for early stopping is val loss when present (distributed.py:187)
                arena.tail[0] = val_loss if val_loss is not None else loss
                loss_buf[i] = loss
                with tracer.phase("exchange_update"):
                    res = comm.reduce_and_update(need_loss)
                if swapped:
                    from sparktorch_b200.ops.module_swap import invalidate_weight_caches

                    invalidate_weight_caches()
            if host_loss is not None:     # per-step device→host read of the step result
                host_loss[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
                d2h_bytes = 4
            window.after_step(i)
            if cfg.verbose:
                print(f"Partition: {rank}. Iteration: {i}. Distributed Loss: {res.loss} "
                      f"Partition Training Loss: {float(loss)}, "
                      f"Partition Validation Loss: {None if val_loss is None else float(val_loss)}")
            if res.stop:
                stopped_at = i
                break
            if cfg.checkpoint_dir and cfg.checkpoint_every > 0 and (i + 1) % cfg.checkpoint_every == 0:
                comm.finish()
                ckpt.save(cfg.checkpoint_dir, rank, {
                    "numel": arena.numel, "flat_p": arena.flat_p.detach().cpu(), "comm": comm.export_state(),
                    "comm_kind": type(comm).__name__, "iters_done": i + 1, "losses": loss_buf[: i + 1].detach().cpu().tolist(),
                    "sampler_state": None if sampler is None else sampler.gen.get_state()})
        comm.finish()
    finally:
        if stall_guard is not None:
            stall_guard.__exit__()
    if tracer.enabled:
        tracer.dump_chrome_trace(trace_path(rank), rank)
    n_done = cfg.iters if stopped_at is None else stopped_at + 1
    losses = loss_buf[:n_done].detach().cpu().tolist()
    comm.average_buffers(model)
    t_end = time.time()
    history = {
        "engine": step.name + ("+graph" if graphed is not None else ""),
        "kernel_swap": swap_report,
        "library_kernels": (swap_report is None or bool(swap_report.get("library_gemm_modules")) or bool(swap_report.get("disabled"))) if device.type == "cuda" else None,
        "comm": type(comm).__name__,
        "world_size": world,
        "train_loss": losses,
        "stopped_at": stopped_at,
        "iters_run": n_done,
        "samples_per_step": (cfg.mini_batch if use_mb else n) * world,
        "loop_seconds": t_end - t_loop,
        "window_ms": window.result_ms(),
        "h2d_bytes_per_step": h2d_bytes,
        "d2h_bytes_per_step": d2h_bytes,
        "streaming": bool(stream_host),
        "resumed_from": start_iter,
        "phases_ms": tracer.summary() if tracer.enabled else None,
        "total_seconds": t_end - t_start,
    }
    state = arena.detach_state_dict() if rank == 0 else None
    return FitResult(state, history)
