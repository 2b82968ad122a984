"""Op substitution for arbitrary modules (ops/module_swap.py): Linear / Conv2d / CE / MSE through the sm_100a kernels
under autograd, the graph-replayed eager step, and the absence of library GEMMs on that path."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, rtol=3e-2, atol=None):
    atol = atol if atol is not None else 3e-2 * float(b.abs().max()) + 1e-6
    torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("shape", [(64, 40, 24), (300, 70, 13), (7, 9, 5)])
def test_linear_forward_backward_matches_fp32(shape):
    from sparktorch_b200.ops.module_swap import swap_modules

    M, K, N = shape
    torch.manual_seed(0)
    lin = nn.Linear(K, N).to(DEV)
    ref = copy.deepcopy(lin)
    lin.weight.grad = torch.zeros_like(lin.weight)          # arena-style pre-allocated gradient: direct accumulation
    swap_modules(lin, None)
    x = torch.randn(3, M, K, device=DEV, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = lin(x), ref(xr)
    _close(y, yr)
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    _close(x.grad, xr.grad)
    _close(lin.weight.grad, ref.weight.grad)
    _close(lin.bias.grad, ref.bias.grad)


@pytest.mark.parametrize("cfg", [(1, 16, 5, 1, 0, 28), (16, 32, 3, 1, 0, 24), (3, 8, 3, 2, 1, 17)])
def test_conv2d_forward_backward_matches_fp32(cfg):
    from sparktorch_b200.ops.module_swap import swap_modules

    cin, cout, k, s, p, hw = cfg
    torch.manual_seed(1)
    conv = nn.Conv2d(cin, cout, k, stride=s, padding=p).to(DEV)
    ref = copy.deepcopy(conv)
    swap_modules(conv, None)
    x = torch.randn(4, cin, hw, hw, device=DEV, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = conv(x), ref(xr)
    assert y.shape == yr.shape
    _close(y, yr)
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    _close(x.grad, xr.grad)
    _close(conv.weight.grad, ref.weight.grad)
    _close(conv.bias.grad, ref.bias.grad)


def test_fused_losses_match_torch():
    from sparktorch_b200.ops.module_swap import swap_modules

    torch.manual_seed(2)
    z = torch.randn(50, 7, device=DEV, requires_grad=True)
    zr = z.detach().clone().requires_grad_(True)
    y = torch.randint(0, 7, (50,), device=DEV)
    _, ce, rep = swap_modules(nn.Identity(), nn.CrossEntropyLoss())
    assert rep["criterion"] == "softmax_ce"
    a, b = ce(z, y), nn.CrossEntropyLoss()(zr, y)
    a.backward()
    b.backward()
    _close(a, b, rtol=1e-4, atol=1e-5)
    _close(z.grad, zr.grad, rtol=2e-2, atol=1e-4)
    t = torch.randn(50, 7, device=DEV)
    _, mse, rep = swap_modules(nn.Identity(), nn.MSELoss())
    z.grad = zr.grad = None
    a, b = mse(z, t), nn.MSELoss()(zr, t)
    a.backward()
    b.backward()
    _close(a, b, rtol=1e-4, atol=1e-5)
    _close(z.grad, zr.grad, rtol=2e-2, atol=1e-4)
    # label smoothing is not substituted
    assert swap_modules(nn.Identity(), nn.CrossEntropyLoss(label_smoothing=0.1))[2]["criterion"] is None


class TanhSoftmaxNet(nn.Module):
    """does not lower to a fused plan: Tanh, a Softmax head (README.md:54 style net of the reference)"""

    def __init__(self):
        super().__init__()
        self.a, self.b, self.c = nn.Linear(64, 96), nn.Linear(96, 48), nn.Linear(48, 6)

    def forward(self, x):
        return torch.softmax(self.c(torch.relu(self.b(torch.tanh(self.a(x))))), dim=1)


def _cls_data(n=1024, d=64, k=6, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)).astype(np.float32)
    W = rng.normal(size=(d, k)).astype(np.float32)
    return X, (X @ W).argmax(1).astype(np.float32).reshape(n, 1)


def test_non_lowerable_net_trains_on_our_kernels_as_one_graph():
    from sparktorch_b200.runtime.engine import TrainConfig, train_partition

    X, Y = _cls_data()
    torch.manual_seed(3)
    net = TanhSoftmaxNet()
    dev = torch.device("cuda", 0)
    kw = {"lr": 5e-3}
    ref = train_partition(0, 1, copy.deepcopy(net), nn.CrossEntropyLoss(), torch.optim.Adam, kw, X, Y,
                          TrainConfig(device="cuda", engine="eager_plain", precision="fp32", iters=60, mini_batch=256, seed=0), dev)
    got = train_partition(0, 1, net, nn.CrossEntropyLoss(), torch.optim.Adam, kw, X, Y,
                          TrainConfig(device="cuda", engine="auto", iters=60, mini_batch=256, seed=0), dev)
    h = got.history
    assert h["engine"] == "eager+graph", h["engine"]
    assert h["kernel_swap"]["linear"] == 3 and h["kernel_swap"]["criterion"] == "softmax_ce" and h["library_kernels"] is False
    a, b = np.array(h["train_loss"]), np.array(ref.history["train_loss"])
    assert len(a) == 60 and np.isfinite(a).all()
    np.testing.assert_allclose(a[:5], b[:5], rtol=2e-2)          # same init, same sampler seed: bf16 rounding only
    assert a[-5:].mean() < a[:5].mean() - 0.02 and abs(a[-5:].mean() - b[-5:].mean()) < 0.05
    net.load_state_dict(got.state_dict)                           # fp32 state_dict loads into the user's module


def test_no_library_gemm_kernels_in_a_swapped_step():
    from sparktorch_b200.ops.module_swap import swap_modules

    try:
        from torch.profiler import ProfilerActivity, profile
    except Exception:
        pytest.skip("torch.profiler unavailable")
    torch.manual_seed(4)
    net = TanhSoftmaxNet().to(DEV)
    net, crit, _ = swap_modules(net, nn.CrossEntropyLoss())
    x = torch.randn(256, 64, device=DEV)
    y = torch.randint(0, 6, (256,), device=DEV)
    crit(net(x), y).backward()
    torch.cuda.synchronize()
    try:
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            crit(net(x), y).backward()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()]
    except Exception as e:
        pytest.skip(f"CUPTI profiling unavailable: {e}")
    if not names:
        pytest.skip("no device events recorded")
    ours = [n for n in names if "st::gemm_bf16" in n]
    lib = [n for n in names if any(t in n.lower() for t in ("cublas", "cutlass", "sgemm", "gemv", "cudnn", "xmma", "nvjet"))]
    assert len(ours) >= 8 and not lib, (sorted(set(names)))
