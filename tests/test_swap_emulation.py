"""The kernel-swap autograd Functions (ops/module_swap.py) with the native extension replaced by its fp32 reference
(ops/reference.py: same entry points, same operand-layout flags): the operand choices of every backward - MN-major weight for
dgrad, both operands MN-major for wgrad, accumulation into the parameter's .grad view, im2col column order = torch's weight
layout, col2im by fold, the pitch handling of the padded bf16 / fp32 buffers - are checked against autograd on the CPU.
The GPU suite (tests/test_gpu_swap.py) checks the same Functions on the real kernels."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from sparktorch_b200.ops import module_swap, native, reference


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(native, "load", lambda required=False: reference)
    yield


def _close(a, b, tol=3e-2):
    torch.testing.assert_close(a, b, rtol=tol, atol=tol)


@pytest.mark.parametrize("shape,bias", [((7, 20), True), ((3, 5, 12), False)])
def test_linear_function(emulated, shape, bias):
    torch.manual_seed(0)
    lin = nn.Linear(shape[-1], 9, bias=bias)
    x = torch.randn(*shape, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    y = module_swap._LinearFn.apply(x, lin.weight, lin.bias, module_swap._WeightCache())
    yr = F.linear(xr, lin.weight.detach(), lin.bias.detach() if bias else None)
    _close(y, yr)
    g = torch.randn_like(yr)
    wr = lin.weight.detach().clone().requires_grad_(True)
    F.linear(xr, wr, lin.bias.detach() if bias else None).backward(g)
    y.backward(g)
    _close(x.grad, xr.grad)
    _close(lin.weight.grad, wr.grad)
    if bias:
        _close(lin.bias.grad, g.reshape(-1, 9).sum(0))


def test_linear_function_accumulates_into_an_existing_gradient_view(emulated):
    """The engine points every parameter's .grad at its slice of the flat arena: wgrad must ADD into it, not replace it."""
    torch.manual_seed(1)
    lin = nn.Linear(8, 4)
    arena = torch.zeros(8 * 4 + 4)
    lin.weight.grad = arena[:32].view(4, 8)
    lin.weight.grad.fill_(0.5)
    x = torch.randn(6, 8)
    y = module_swap._LinearFn.apply(x, lin.weight, lin.bias, module_swap._WeightCache())
    g = torch.randn(6, 4)
    y.backward(g)
    want = 0.5 + g.t() @ x
    _close(arena[:32].view(4, 8), want)
    assert lin.weight.grad.data_ptr() == arena.data_ptr()


@pytest.mark.parametrize("stride,pad,k", [(1, 1, 3), (2, 1, 3), (1, 0, 5)])
def test_conv_function(emulated, stride, pad, k):
    torch.manual_seed(2)
    conv = nn.Conv2d(3, 6, k, stride=stride, padding=pad)
    x = (torch.randn(2, 3, 11, 9) * 0.5).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    y = module_swap._Conv2dFn.apply(x, conv.weight, conv.bias, module_swap._WeightCache(), stride, pad)
    yr = F.conv2d(xr, wr, conv.bias.detach(), stride=stride, padding=pad)
    assert y.shape == yr.shape
    _close(y, yr)
    g = torch.randn_like(yr) * 0.2
    yr.backward(g)
    y.backward(g)
    _close(x.grad, xr.grad)
    _close(conv.weight.grad, wr.grad)
    _close(conv.bias.grad, g.sum(dim=(0, 2, 3)))


def test_loss_functions(emulated):
    torch.manual_seed(3)
    z = torch.randn(12, 10, requires_grad=True)
    t = torch.randint(0, 10, (12,))
    zr = z.detach().clone().requires_grad_(True)
    loss = module_swap._SoftmaxCEFn.apply(z, t)
    ref = F.cross_entropy(zr, t)
    _close(loss, ref, 1e-3)
    (2.0 * loss).backward()
    (2.0 * ref).backward()
    _close(z.grad, zr.grad, 1e-2)                      # dlogits travel as bf16
    p = torch.randn(9, 3, requires_grad=True)
    tgt = torch.randn(9, 3)
    pr = p.detach().clone().requires_grad_(True)
    loss = module_swap._MSEFn.apply(p, tgt)
    ref = F.mse_loss(pr, tgt)
    _close(loss, ref, 1e-3)
    loss.backward()
    ref.backward()
    _close(p.grad, pr.grad, 1e-2)
