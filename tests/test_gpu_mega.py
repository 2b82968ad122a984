"""K-STEP (csrc/mlp_step.cu): whole training steps in one persistent kernel, against (a) the round-1 multi-kernel
graph of the same plan (SPARKTORCH_MEGA=0: same kernels' math, 9 launches per step) and (b) fp32 eager autograd."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _data(n=512, d=784, classes=10, seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.integers(0, 256, size=(n, d)) * (rng.random((n, d)) < 0.2)).astype(np.float32) / 64.0
    W = rng.normal(size=(d, classes)).astype(np.float32)
    Y = (X @ W).argmax(1).astype(np.float32).reshape(n, 1)
    return X, Y


def _fit(model, crit, opt, kw, X, Y, mega=True, **cfgkw):
    from sparktorch_b200.runtime.engine import TrainConfig, train_partition

    os.environ["SPARKTORCH_MEGA"] = "1" if mega else "0"
    try:
        cfgkw.setdefault("seed", 0)
        return train_partition(0, 1, model, crit, opt, kw, X, Y, TrainConfig(device="cuda", engine="fused", **cfgkw),
                               torch.device("cuda", 0))
    finally:
        os.environ.pop("SPARKTORCH_MEGA", None)


def test_one_step_matches_the_multi_kernel_graph():
    from sparktorch_b200.models import mnist_dnn

    X, Y = _data(n=256)
    torch.manual_seed(0)
    net = mnist_dnn()
    a = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.SGD, {"lr": 0.05}, X, Y, mega=True, iters=1)
    b = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.SGD, {"lr": 0.05}, X, Y, mega=False, iters=1)
    assert a.history["step_kernel"] is not None and b.history["step_kernel"] is None
    assert a.history["gpu_launches_per_step"] == 1 and b.history["gpu_launches_per_step"] == 9
    np.testing.assert_allclose(a.history["train_loss"], b.history["train_loss"], rtol=1e-4)
    for k in a.state_dict:       # SGD: parameter difference = lr * gradient difference (fp32 accumulation order only)
        torch.testing.assert_close(a.state_dict[k], b.state_dict[k], rtol=1e-3, atol=2e-5, msg=lambda m, k=k: f"{k}: {m}")


@pytest.mark.parametrize("steps_per_launch", ["1", "5", "64"])
def test_many_steps_adam_match_the_multi_kernel_graph(steps_per_launch):
    from sparktorch_b200.models import mnist_dnn

    X, Y = _data(n=2048)
    torch.manual_seed(1)
    net = mnist_dnn()
    os.environ["SPARKTORCH_MEGA_STEPS"] = steps_per_launch
    try:
        a = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, mega=True, iters=40,
                 mini_batch=256, sampling="sweep")
    finally:
        os.environ.pop("SPARKTORCH_MEGA_STEPS", None)
    b = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, mega=False, iters=40,
             mini_batch=256, sampling="sweep")
    assert a.history["iters_run"] == 40 and a.history["step_kernel"]["steps_per_launch"] == int(steps_per_launch)
    np.testing.assert_allclose(a.history["train_loss"], b.history["train_loss"], rtol=2e-2, atol=2e-3)
    assert a.history["train_loss"][-1] < 0.8 * a.history["train_loss"][0]
    for k in a.state_dict:
        diff = (a.state_dict[k] - b.state_dict[k]).abs()
        assert (diff < 5e-3).float().mean().item() > 0.98, k


def test_tracks_eager_fp32_with_odd_shapes_bias_free_layers_and_mse():
    """B = 300 (3 m-tiles, ragged last one), widths that are no multiple of the tile sizes, a layer without bias, a linear
    layer without ReLU, MSE head with 7 outputs, SGD momentum."""
    from sparktorch_b200.runtime.engine import TrainConfig, train_partition

    rng = np.random.default_rng(4)
    X = rng.normal(size=(300, 70)).astype(np.float32)
    Y = np.tanh(X[:, :7] * 0.5 + X[:, 7:14]).astype(np.float32)
    torch.manual_seed(2)
    net = nn.Sequential(nn.Linear(70, 200), nn.ReLU(), nn.Linear(200, 136, bias=False), nn.Linear(136, 24), nn.ReLU(), nn.Linear(24, 7))
    kw = {"lr": 0.02, "momentum": 0.9}
    ref = train_partition(0, 1, copy.deepcopy(net), nn.MSELoss(), torch.optim.SGD, kw, X, Y,
                          TrainConfig(device="cuda", engine="eager", comm="collective", precision="fp32", iters=25, seed=0),
                          torch.device("cuda", 0))
    got = _fit(net, nn.MSELoss(), torch.optim.SGD, kw, X, Y, iters=25)
    assert got.history["step_kernel"] is not None
    a, b = np.array(got.history["train_loss"]), np.array(ref.history["train_loss"])
    np.testing.assert_allclose(a, b, rtol=0.05, atol=5e-3)
    assert a[-1] < 0.95 * a[0]
    for k in ref.state_dict:
        torch.testing.assert_close(got.state_dict[k], ref.state_dict[k], rtol=0.1, atol=2e-2, msg=lambda m, k=k: f"{k}: {m}")


def test_early_stop_random_sampling_and_streaming_paths():
    from sparktorch_b200.models import mnist_dnn
    from tests.nets import Regressor

    rng = np.random.default_rng(3)
    X = rng.normal(size=(300, 10)).astype(np.float32)
    Y = rng.normal(size=(300, 1)).astype(np.float32)
    res = _fit(Regressor(), nn.MSELoss(), torch.optim.SGD, {"lr": 0.05, "maximize": True}, X, Y, iters=50, early_stop_patience=3)
    assert res.history["step_kernel"] is not None and res.history["stopped_at"] == 3 and res.history["iters_run"] == 4
    X, Y = _data(n=2048)
    torch.manual_seed(1)
    net = mnist_dnn()
    a = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, iters=24, mini_batch=256, sampling="sweep")
    b = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, iters=24, mini_batch=256, streaming=True)
    c = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, iters=24, mini_batch=256, sampling="random")
    assert b.history["streaming"] and b.history["step_kernel"]["steps_per_launch"] == 4      # SPARKTORCH_STREAM_STEPS
    np.testing.assert_allclose(a.history["train_loss"], b.history["train_loss"], rtol=1e-3, atol=1e-4)
    assert c.history["gpu_launches_per_step"] == 2 and c.history["train_loss"][-1] < c.history["train_loss"][0]


def test_checkpoint_resume_is_bit_exact_on_the_fused_path(tmp_path):
    """checkpointDir keeps the K-STEP engine (round 1 fell back to the eager engine): 30 straight steps == 20 steps, a new
    process-equivalent fit that resumes from the checkpoint, 10 more - bit for bit (deterministic reductions, the step
    kernel takes iteration / optimizer-step numbers as launch parameters)."""
    from sparktorch_b200.models import mnist_dnn

    X, Y = _data(n=2048)
    torch.manual_seed(5)
    net = mnist_dnn()
    kw = dict(mini_batch=256, sampling="sweep")
    straight = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, iters=30, **kw)
    d = str(tmp_path / "ck")
    first = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, iters=20, checkpoint_dir=d,
                 checkpoint_every=10, **kw)
    assert first.history["step_kernel"] is not None and first.history["iters_run"] == 20
    resumed = _fit(copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 1e-3}, X, Y, iters=30, checkpoint_dir=d,
                   checkpoint_every=10, resume=True, **kw)
    assert resumed.history["resumed_from"] == 20 and resumed.history["iters_run"] == 30 and resumed.history["step_kernel"] is not None
    assert resumed.history["train_loss"] == straight.history["train_loss"]
    for k in straight.state_dict:
        assert torch.equal(resumed.state_dict[k], straight.state_dict[k]), k
