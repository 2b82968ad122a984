"""Auxiliary subsystems the reference lacks: checkpoint/resume, failure
handling (barrier-stage retry), tracing, metrics."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import sparktorch_b200  # noqa: F401
from sparktorch_b200 import SparkTorch, serialize_torch_obj
from sparktorch_b200.runtime.engine import TrainConfig, train_partition
from sparktorch_b200.utils import metrics
from tests.nets import Regressor


def _xy(n=200, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 10)).astype(np.float32)
    Y = (X[:, :1] - 0.5 * X[:, 1:2]).astype(np.float32)
    return X, Y


def test_checkpoint_resume_is_bit_exact(tmp_path):
    X, Y = _xy()
    torch.manual_seed(0)
    net = Regressor()
    import copy

    dev = torch.device("cpu")
    full = train_partition(0, 1, copy.deepcopy(net), nn.MSELoss(), torch.optim.Adam, {"lr": 1e-2}, X, Y,
                           TrainConfig(iters=20, seed=1), dev)
    d = str(tmp_path / "ck")
    part = train_partition(0, 1, copy.deepcopy(net), nn.MSELoss(), torch.optim.Adam, {"lr": 1e-2}, X, Y,
                           TrainConfig(iters=12, seed=1, checkpoint_dir=d, checkpoint_every=6), dev)
    assert os.path.exists(os.path.join(d, "rank0.pt")) and part.history["iters_run"] == 12
    rest = train_partition(0, 1, copy.deepcopy(net), nn.MSELoss(), torch.optim.Adam, {"lr": 1e-2}, X, Y,
                           TrainConfig(iters=20, seed=1, checkpoint_dir=d, checkpoint_every=6, resume=True), dev)
    assert rest.history["resumed_from"] == 12
    np.testing.assert_allclose(rest.history["train_loss"], full.history["train_loss"], rtol=0, atol=0)
    for k in full.state_dict:
        assert torch.equal(rest.state_dict[k], full.state_dict[k])


def test_barrier_stage_is_retried_and_resumes_after_a_task_failure(tmp_path, spark, monkeypatch):
    rng = np.random.default_rng(0)
    X = rng.normal(size=(80, 10))
    y = (X[:, 0] > 0).astype(np.float64)
    df = spark.createDataFrame({"label": y, "features": X}, numPartitions=2)
    obj = serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.Adam, lr=0.01)
    d = str(tmp_path / "ck")
    monkeypatch.setenv("SPARKTORCH_FAULT_INJECT", "1:7")      # executor rank 1 dies at iteration 7, once
    est = SparkTorch(inputCol="features", labelCol="label", predictionCol="p", torchObj=obj, iters=12, seed=3,
                     checkpointDir=d, checkpointEvery=5)
    # the executor process inherits the env at spawn time; make sure a fresh pool sees it
    spark.sparkContext._cluster.shutdown()
    spark.sparkContext._cluster.__init__(2)
    stm = est.fit(df)
    h = stm.getTrainingHistory()
    assert h["resumed_from"] == 5 and h["iters_run"] == 12 and len(h["train_loss"]) == 12
    assert os.path.exists(os.path.join(d, ".fault_1_7"))


def test_tracing_and_metrics(tmp_path, monkeypatch):
    X, Y = _xy()
    monkeypatch.setenv("SPARKTORCH_TRACE", str(tmp_path / "trace.json"))
    res = train_partition(0, 1, Regressor(), nn.MSELoss(), torch.optim.SGD, {"lr": 0.05}, X, Y,
                          TrainConfig(iters=5, seed=0), torch.device("cpu"))
    tr = json.load(open(tmp_path / "trace.rank0.json"))
    names = {e["name"] for e in tr["traceEvents"]}
    assert {"forward_backward", "exchange_update"} <= names
    assert res.history["phases_ms"]["forward_backward"]["count"] == 5
    text = metrics.to_prometheus(res.history)
    assert "sparktorch_iterations_total" in text and metrics.throughput(res.history) > 0
    assert len(metrics.to_jsonl(res.history).splitlines()) == 6


def test_progress_watchdog_fires_on_stall_and_stays_quiet_on_progress():
    import time

    from sparktorch_b200.utils.watchdog import ProgressWatchdog

    fired = []
    counter = [0]
    with ProgressWatchdog(lambda: counter[0], timeout_s=0.3, label="t", on_stall=fired.append, poll_s=0.05) as wd:
        for _ in range(12):              # steady progress for 0.6 s: must not fire
            counter[0] += 1
            time.sleep(0.05)
        assert not fired and not wd.fired
        time.sleep(0.6)                  # progress stops
    assert len(fired) == 1 and "no progress" in fired[0] and "iteration 12" in fired[0]
    quiet = []
    with ProgressWatchdog(lambda: 0, timeout_s=0, on_stall=quiet.append):    # disabled
        time.sleep(0.1)
    assert not quiet


def test_bench_contract_on_cpu():
    """bench.py prints ONE JSON line with the keys the driver parses (CPU run of the eager engine)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--device", "cpu", "--steps", "3", "--warmup", "3",
                          "--rows-per-gpu", "1024"], capture_output=True, text=True, timeout=600, cwd=root)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches"):
        assert k in r, k
    assert r["n_gpus"] == 1 and r["steps"] == 3 and r["warmup"] == 3 and r["higher_is_better"] is True and r["scaling"] == "weak"
    assert r["value"] > 0 and abs(r["value"] - 256 * 3 / (r["ms_per_step"] * 3 / 1e3)) / r["value"] < 1e-6
    for k in ("model", "global_batch", "seq_len", "parallelism"):
        assert k in r["config"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in r["e2e"], k
    assert r["config"]["l2"].startswith("WARNING")          # 1024 rows do not exceed L2 and the line says so


def test_batch_predictor_bounds_staging_by_bytes():
    import numpy as np
    import torch
    import torch.nn as nn

    from sparktorch_b200.runtime.predict import BatchPredictor

    calls = []

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(1 << 20, 2)

        def forward(self, x):
            calls.append(x.shape[0])
            return self.fc(x)

    bp = BatchPredictor(Net(), torch.device("cpu"), batch_size=65536)
    out = bp.predict_block(np.zeros((130, 1 << 20), dtype=np.float32))       # 4 MiB rows -> 64 rows per transfer
    assert out.shape == (130, 2) and calls == [64, 64, 2]


# ------------------------------------------------------------------------------------------------ round-2 advisor items
def test_one_hot_labels_with_cross_entropy_and_minibatch_train_on_the_right_targets():
    """k-column (one-hot / probability) labels stay float class probabilities; only a single label column becomes class
    indices (the reference's first criterion(...) try succeeds for probability targets, distributed.py:152-157)."""
    import numpy as np
    import torch
    import torch.nn as nn

    from sparktorch_b200.runtime.engine import TrainConfig, prepare_targets, train_partition

    y1 = prepare_targets(torch.tensor([[1.0], [0.0], [2.0]]), nn.CrossEntropyLoss())
    assert y1.dtype == torch.int64 and y1.shape == (3,)
    yk = prepare_targets(torch.eye(3), nn.CrossEntropyLoss())
    assert yk.dtype == torch.float32 and yk.shape == (3, 3)
    rng = np.random.default_rng(0)
    X = rng.normal(size=(96, 6)).astype(np.float32)
    cls = (X[:, 0] > 0).astype(np.int64) + (X[:, 1] > 0).astype(np.int64)
    Y = np.eye(3, dtype=np.float32)[cls]
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 16), nn.ReLU(), nn.Linear(16, 3))
    res = train_partition(0, 1, net, nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 2e-2}, X, Y,
                          TrainConfig(iters=150, mini_batch=16, seed=0), torch.device("cpu"))
    loss = res.history["train_loss"]
    assert np.mean(loss[-10:]) < 0.6 * np.mean(loss[:10])
    with torch.no_grad():
        net.load_state_dict(res.state_dict)
        acc = (net(torch.from_numpy(X)).argmax(1).numpy() == cls).mean()
    assert acc > 0.85


def test_checkpoint_generations_and_cross_rank_agreement(tmp_path):
    from sparktorch_b200.utils import checkpoint as ckpt

    d = str(tmp_path)
    ckpt.save(d, 0, {"iters_done": 4, "numel": 8})
    ckpt.save(d, 0, {"iters_done": 8, "numel": 8})
    ckpt.save(d, 0, {"iters_done": 12, "numel": 8})
    gens = ckpt.load_generations(d, 0)
    assert sorted(gens) == [8, 12]                     # two generations are kept
    assert ckpt.agree_on_iteration(sorted(gens), 1, None) == 12
    assert ckpt.agree_on_iteration([], 1, None) == 0


def _resume_worker(rank, world, port, d, out):
    import os

    import torch
    import torch.distributed as dist

    from sparktorch_b200.utils import checkpoint as ckpt

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 crashed before writing generation 12: it holds {4, 8}, rank 1 holds {8, 12}
    out[rank] = ckpt.agree_on_iteration(sorted(ckpt.load_generations(d, rank)), world, torch.device("cpu"))
    dist.destroy_process_group()


def test_resume_picks_the_newest_iteration_every_rank_holds(tmp_path):
    import socket

    import torch.multiprocessing as mp

    from sparktorch_b200.utils import checkpoint as ckpt

    d = str(tmp_path)
    for it in (4, 8):
        ckpt.save(d, 0, {"iters_done": it, "numel": 8})
    for it in (8, 12):
        ckpt.save(d, 1, {"iters_done": it, "numel": 8})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_resume_worker, args=(2, port, d, out), nprocs=2, join=True)
    assert dict(out) == {0: 8, 1: 8}


def test_asynchrony_compensation_keeps_the_total_momentum():
    import torch

    from sparktorch_b200.parallel.hogwild_p2p import compensate_asynchrony
    from sparktorch_b200.runtime.optim import parse_optimizer

    adam = parse_optimizer(torch.optim.Adam, {"lr": 1e-3})
    assert compensate_asynchrony(adam, 1) is adam
    assert abs(compensate_asynchrony(adam, 2).beta1 - 0.8) < 1e-9          # 1 - beta' = W (1 - beta)
    assert abs(compensate_asynchrony(adam, 8).beta1 - 0.2) < 1e-9
    assert compensate_asynchrony(adam, 16).beta1 == 0.0                     # clipped: asynchrony alone is already more momentum
    sgd = parse_optimizer(torch.optim.SGD, {"lr": 0.1, "momentum": 0.9})
    assert abs(compensate_asynchrony(sgd, 4).momentum - 0.6) < 1e-9
    plain = parse_optimizer(torch.optim.SGD, {"lr": 0.1})
    assert compensate_asynchrony(plain, 8).momentum == 0.0


def test_module_swap_is_a_no_op_without_the_native_extension_or_in_fp32():
    import torch
    import torch.nn as nn

    from sparktorch_b200.ops.module_swap import swap_modules, unswap_modules

    net = nn.Sequential(nn.Linear(4, 3), nn.Tanh(), nn.Linear(3, 2))
    x = torch.randn(5, 4)
    want = net(x)
    _, crit, rep = swap_modules(net, nn.CrossEntropyLoss(), precision="fp32")
    assert rep["disabled"] == "precision=fp32" and rep["linear"] == 0 and len(rep["library_gemm_modules"]) == 2
    assert isinstance(crit, nn.CrossEntropyLoss)
    net2, crit2, rep2 = swap_modules(net, nn.CrossEntropyLoss())             # CPU box: extension present, tensors on the CPU
    torch.testing.assert_close(net2(x), want)                                # CPU inputs take the stock path
    loss = crit2(net2(x), torch.tensor([0, 1, 0, 1, 1]))
    loss.backward()
    assert all(p.grad is not None for p in net.parameters())
    unswap_modules(net)
    assert "forward" not in vars(net[0])
