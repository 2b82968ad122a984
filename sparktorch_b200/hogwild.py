"""Asynchronous parameter-server training (``mode='hogwild'``).

Surface of the reference module (sparktorch/hogwild.py:31-186): HTTP helpers,
``handle_model`` (worker) and ``train`` (driver).  The worker loop is new:

* pulls/pushes the flat fp32 arena as raw bytes over a keep-alive connection
  (``/flat/*`` routes) instead of dill'ed dicts/lists;
* zeroes gradients every iteration (the reference accumulates them forever,
  quirk Q9);
* honours ``device`` (reference workers are CPU only, quirk Q10).

On a single B200 node ``train`` switches to the GPU-resident owner-sharded
variant (parallel/hogwild_p2p.py): no HTTP, parameters are pulled with a P2P
all-gather fused into the first GEMM and gradients are pushed with peer stores.
"""
from __future__ import annotations

import http.client
import json
from typing import Dict, List

import dill
import numpy as np
import torch

from sparktorch_b200._spark import ensure_pyspark

ensure_pyspark()

from pyspark.rdd import RDD  # noqa: E402

from sparktorch_b200.runtime.engine import EagerStep, DeviceSampler, prepare_targets, resolve_device  # noqa: E402
from sparktorch_b200.runtime.flat import FlatArena  # noqa: E402
from sparktorch_b200.server import Server  # noqa: E402
from sparktorch_b200.util import DataObj, load_base_torch, load_torch_model, rows_to_block, split_validation  # noqa: E402


# ------------------------------------------------------------------ HTTP client
class _PSClient:
    """Keep-alive client with one reconnect+retry per call and a 10 s timeout
    (retry policy of reference hogwild.py:31-57)."""

    def __init__(self, master_url: str):
        host, _, port = master_url.partition(":")
        self.host, self.port = host, int(port or 80)
        self.conn = None

    def _request(self, method: str, path: str, body: bytes = b"", ctype: str = "application/octet-stream") -> bytes:
        last = None
        for _attempt in range(2):
            try:
                if self.conn is None:
                    self.conn = http.client.HTTPConnection(self.host, self.port, timeout=10)
                self.conn.request(method, path, body=body, headers={"Content-Type": ctype})
                resp = self.conn.getresponse()
                data = resp.read()
                if resp.status != 200:
                    raise RuntimeError(f"{method} {path} -> {resp.status}")
                return data
            except Exception as e:
                last = e
                try:
                    if self.conn is not None:
                        self.conn.close()
                finally:
                    self.conn = None
        raise last  # type: ignore[misc]

    def get(self, path: str) -> bytes:
        return self._request("GET", path)

    def post(self, path: str, body: bytes, ctype: str = "application/octet-stream") -> bytes:
        return self._request("POST", path, body, ctype)


def get_state_dict(master_url: str = "localhost:3000", retry=True) -> Dict:
    return dill.loads(_PSClient(master_url).get("/parameters"))


def put_deltas_to_server(delta: List, master_url: str = "localhost:3000", retry=True):
    try:
        _PSClient(master_url).post("/update", dill.dumps(delta))
    except Exception:
        pass  # dropped pushes are tolerated (reference hogwild.py:44-49)


def put_early_stop(loss, master_url: str = "localhost:3000", retry=True):
    return json.loads(_PSClient(master_url).post("/losses", json.dumps({"loss": loss}).encode(), "application/json"))


def get_main(master_url: str = "localhost:3000") -> str:
    return _PSClient(master_url).get("/").decode()


# ------------------------------------------------------------------ worker
def handle_model(
    data: List[DataObj],
    torch_obj: str,
    master_url: str = "localhost:3000",
    iters: int = 1000,
    verbose: int = 1,
    early_stop_patience: int = -1,
    mini_batch: int = -1,
    validation_pct: float = 0,
    device: str = "cpu",
    seed: int = None,
):
    if data is None:
        return "finished"
    X, Y = rows_to_block(data)
    if X is None:
        return "finished"
    try:
        from pyspark.taskcontext import TaskContext

        ctx = TaskContext.get()
        pid = ctx.partitionId() if ctx is not None else 0
    except Exception:
        pid = 0
    dev = resolve_device(device, pid)
    raw = load_torch_model(torch_obj, build_optimizer=False)
    model, criterion = raw.model.to(dev), raw.criterion
    model.train()
    rng = np.random.default_rng(None if seed is None else seed + pid)
    xt, yt = torch.from_numpy(X), (torch.from_numpy(Y) if Y is not None else None)
    x_val = y_val = None
    if validation_pct > 0:
        tr, va = split_validation(X.shape[0], validation_pct, rng)
        if len(va):
            x_val = xt[torch.from_numpy(va)].to(dev)
            y_val = prepare_targets(yt[torch.from_numpy(va)].to(dev), criterion) if yt is not None else x_val
            xt = xt[torch.from_numpy(tr)]
            yt = yt[torch.from_numpy(tr)] if yt is not None else None
    x_all = xt.to(dev)
    y_all = prepare_targets(yt.to(dev), criterion) if yt is not None else x_all

    arena = FlatArena(model, dev)
    step = EagerStep(model, criterion, arena)
    client = _PSClient(master_url)
    n = x_all.shape[0]
    sampler = DeviceSampler(n, mini_batch, dev, seed) if 0 < mini_batch < n else None
    host_p = torch.empty(arena.numel, dtype=torch.float32, pin_memory=dev.type == "cuda")
    for i in range(iters):
        if sampler is not None:
            idx = sampler.next()
            xb = x_all[idx]
            yb = xb if y_all is x_all else y_all[idx]
        else:
            xb, yb = x_all, y_all
        blob = client.get("/flat/parameters")
        host_p.copy_(torch.from_numpy(np.frombuffer(bytearray(blob), dtype=np.float32, offset=8, count=arena.numel)))
        arena.flat_p.copy_(host_p, non_blocking=True)
        loss = step.run(xb, yb)                       # zeroes grads first (quirk Q9 fixed)
        val_loss = step.evaluate(x_val, y_val) if x_val is not None else None
        client.post("/flat/update", arena.grads.detach().cpu().numpy().tobytes())
        loss_v = float(loss)
        val_v = None if val_loss is None else float(val_loss)
        if verbose:
            print(f"Partition: {pid}. Iteration: {i}. Loss: {loss_v}, Val Loss: {val_v}")
        if early_stop_patience > 0:
            use = val_v if val_v is not None else loss_v
            reply = json.loads(client.post("/losses", json.dumps({"loss": use}).encode(), "application/json"))
            if reply["stop"]:
                break
    return "finished"


# ------------------------------------------------------------------ GPU-resident variant
def handle_model_gpu(index, data, torch_obj, iters=1000, verbose=1, early_stop_patience=-1, mini_batch=-1,
                     validation_pct=0.0, device="cuda", seed=None, acquire_lock=True, window_len=4, time_window=None,
                     sampling="random"):
    """Barrier-task body of ``mode='hogwild', device='cuda'``: no HTTP, the GPUs' own
    memory is the parameter server (parallel/hogwild_p2p.py)."""
    from pyspark.taskcontext import BarrierTaskContext

    from sparktorch_b200.parallel import rendezvous
    from sparktorch_b200.parallel.hogwild_p2p import train_partition_hogwild
    from sparktorch_b200.runtime.engine import TrainConfig

    context = BarrierTaskContext.get()
    dev = resolve_device(device, context.partitionId())
    torch.cuda.set_device(dev)
    rank, world = rendezvous.init_from_barrier(context, dev)
    try:
        raw = load_torch_model(torch_obj, build_optimizer=False)
        X, Y = rows_to_block(data)
        if X is None:
            raise RuntimeError(f"partition {index} is empty")
        cfg = TrainConfig(iters=iters, mini_batch=mini_batch, validation_pct=validation_pct,
                          early_stop_patience=early_stop_patience, verbose=verbose, device=str(dev), seed=seed,
                          time_window=tuple(time_window) if time_window else None, sampling=sampling)
        # the graph-captured worker with K-AG forward GEMMs - with acquireLock (the default) its remote-apply push takes the
        # owners' stripe locks, without it the push is lock-free; models that do not lower fall through to the eager worker
        from sparktorch_b200.parallel.hogwild_p2p import train_partition_hogwild_fused

        res = train_partition_hogwild_fused(rank, world, raw.model, raw.criterion, raw.optimizer, raw.optimizer_params,
                                            X, Y, cfg, dev, window_len, acquire_lock=acquire_lock)
        if res is None:
            res = train_partition_hogwild(rank, world, raw.model, raw.criterion, raw.optimizer, raw.optimizer_params, X, Y,
                                          cfg, dev, acquire_lock, window_len)
    finally:
        rendezvous.shutdown()
    return [{"state_dict": res.state_dict, "history": res.history}] if rank == 0 else []


def train_gpu(rdd, torch_obj, iters=10, partition_shuffles=1, verbose=1, early_stop_patience=-1, mini_batch=-1,
              validation_pct=0.0, device="cuda", seed=None, acquire_lock=True, window_len=4, time_window=None,
              sampling="random"):
    from sparktorch_b200.distributed import _reseed_blob, mapPartitionsWithIndex

    blob, _ = load_base_torch(torch_obj)
    state_dict = history = None
    for r in range(partition_shuffles):
        def task(i, x, blob=blob):
            return handle_model_gpu(i, x, blob, iters=iters, verbose=verbose, early_stop_patience=early_stop_patience,
                                    mini_batch=mini_batch, validation_pct=validation_pct, device=device, seed=seed,
                                    acquire_lock=acquire_lock, window_len=window_len, time_window=time_window,
                                    sampling=sampling)

        task._accepts_blocks = True
        out = mapPartitionsWithIndex(rdd, task).collect()
        state_dict, history = out[0]["state_dict"], out[0]["history"]
        if partition_shuffles - r > 1:
            rdd = rdd.repartition(rdd.getNumPartitions())
            blob = _reseed_blob(blob, state_dict)       # weights carry over, like the persistent server
    return state_dict, history


# ------------------------------------------------------------------ driver
def train(
    rdd: RDD,
    torch_obj: str,
    server: Server,
    iters: int = 10,
    partition_shuffles: int = 1,
    verbose: int = 1,
    early_stop_patience: int = -1,
    mini_batch: int = -1,
    validation_pct: float = 0.0,
    device: str = "cpu",
    seed: int = None,
) -> Dict:
    """Run ``partition_shuffles`` rounds of workers against ``server``; returns
    the final state_dict and stops the server (reference hogwild.py:145-186)."""
    try:
        blob, _ = load_base_torch(torch_obj)
        master_url = str(server.master_url)
        for r in range(partition_shuffles):
            def work(x, blob=blob):
                return [handle_model(
                    x, torch_obj=blob, master_url=master_url, iters=iters, verbose=verbose,
                    early_stop_patience=early_stop_patience, mini_batch=mini_batch,
                    validation_pct=validation_pct, device=device, seed=seed)]

            work._accepts_blocks = True
            rdd.mapPartitions(work).foreach(lambda x: x)
            if partition_shuffles - r > 1:
                rdd = rdd.repartition(rdd.getNumPartitions())
        state_dict = server.state_dict()
        server.stop_server()
        return state_dict
    except Exception:
        server.stop_server()
        raise
