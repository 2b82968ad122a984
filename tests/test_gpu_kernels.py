"""Numerics of the hand-written sm_100a kernels against plain PyTorch fp32
references of the same op (run on the B200 box: ``pytest -m gpu``)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from sparktorch_b200.ops import native

    return native.load(required=True)


def dev():
    return torch.device("cuda", 0)


def _pad8(n):
    return (n + 7) // 8 * 8


def _mk(rows, cols, scale=1.0, seed=0):
    """bf16 [rows, cols] matrix living in a buffer with a padded pitch."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    full = torch.zeros(rows, _pad8(cols), dtype=torch.bfloat16, device=dev())
    full[:, :cols] = (torch.randn(rows, cols, generator=g) * scale).to(torch.bfloat16).to(dev())
    return full[:, :cols]


SHAPES = [
    (256, 256, 784),    # MNIST-DNN layer 1 fwd
    (256, 10, 256),     # classifier head (N padded to 16)
    (128, 64, 64),
    (300, 200, 100),    # ragged everything
    (17, 20, 10),       # tiny (API test nets)
    (1024, 512, 2048),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_operand_layouts(C, M, N, K, a_mn, b_mn):
    A = _mk(K, M, seed=1) if a_mn else _mk(M, K, seed=1)
    B = _mk(K, N, seed=2) if b_mn else _mk(N, K, seed=2)
    ref = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    out = torch.full((M, _pad8(N)), float("nan"), dtype=torch.float32, device=dev())[:, :N]
    C.gemm_bf16(A, a_mn, B, b_mn, M, N, K, 1.0, None, False, None, None, None, out, None, False, None, 0, 1)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    tol = 1e-2 * math.sqrt(K) * 0.25 + 1e-3
    assert err < tol, f"max err {err} (tol {tol})"


@pytest.mark.parametrize("block_n", [16, 32, 64, 128, 256])
def test_gemm_block_n_variants(C, block_n):
    M, N, K = 384, 256, 320
    A, B = _mk(M, K, seed=3), _mk(N, K, seed=4)
    out = torch.zeros(M, N, dtype=torch.float32, device=dev())
    C.gemm_bf16(A, False, B, False, M, N, K, 1.0, None, False, None, None, None, out, None, False, None, block_n, 1)
    torch.testing.assert_close(out, A.float() @ B.float().t(), rtol=1e-3, atol=5e-2)


def test_gemm_deep_k_pipeline_wraps(C):
    # K = 64 * 40 blocks > 8 stages: exercises phase flips of the smem ring
    M, N, K = 128, 128, 2560
    A, B = _mk(M, K, 0.5, 5), _mk(N, K, 0.5, 6)
    out = torch.zeros(M, N, dtype=torch.float32, device=dev())
    C.gemm_bf16(A, False, B, False, M, N, K, 1.0, None, False, None, None, None, out, None, False, None, 0, 1)
    torch.testing.assert_close(out, A.float() @ B.float().t(), rtol=2e-3, atol=1e-1)


def test_gemm_fused_forward_epilogue(C):
    # y = relu(x W^T + b) in bf16, plus the transposed copy
    Bt, I, O = 256, 784, 256
    x, w = _mk(Bt, I, 0.1, 7), _mk(O, I, 0.1, 8)
    b = torch.randn(O, device=dev())
    y = torch.zeros(Bt, O, dtype=torch.bfloat16, device=dev())
    yt = torch.zeros(O, Bt, dtype=torch.bfloat16, device=dev())
    C.gemm_bf16(x, False, w, False, Bt, O, I, 1.0, b, True, None, y, yt, None, None, False, None, 0, 1)
    ref = torch.relu(x.float() @ w.float().t() + b)
    torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(yt.float(), ref.t(), rtol=2e-2, atol=2e-2)


def test_gemm_fused_backward_epilogue(C):
    # dz_prev = (dz W) * (a_prev > 0); db_prev = column sums  -- dgrad with W as MN-major B operand
    Bt, I, O = 256, 256, 10
    dz = _mk(Bt, O, 0.1, 9)           # [B, out]
    w = _mk(O, I, 0.1, 10)            # [out, in]  = [K, N] storage -> b_mn
    a_prev = torch.relu(_mk(Bt, I, 1.0, 11))
    out = torch.zeros(Bt, I, dtype=torch.bfloat16, device=dev())
    db = torch.zeros(I, dtype=torch.float32, device=dev())
    C.gemm_bf16(dz, False, w, True, Bt, I, O, 1.0, None, False, a_prev, out, None, None, None, False, db, 0, 1)
    ref = (dz.float() @ w.float()) * (a_prev.float() > 0)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-3)
    torch.testing.assert_close(db, ref.sum(0), rtol=2e-2, atol=2e-2)


def test_gemm_wgrad_split_k_atomic(C):
    # dW[out,in] = dz^T a : both operands MN-major ([B, *] storage), reduction over a big batch, split-K red.add
    Bt, I, O = 4096, 784, 256
    dz, a = _mk(Bt, O, 0.05, 12), _mk(Bt, I, 0.5, 13)
    dw = torch.zeros(O, I, dtype=torch.float32, device=dev())
    C.gemm_bf16(dz, True, a, True, O, I, Bt, 1.0, None, False, None, None, None, dw, None, True, None, 0, 8)
    torch.testing.assert_close(dw, dz.float().t() @ a.float(), rtol=2e-3, atol=5e-2)
    # transposed fp32 store: dW^T computed as a^T dz and written back as dW
    dw2 = torch.zeros(O, I, dtype=torch.float32, device=dev())
    C.gemm_bf16(a, True, dz, True, I, O, Bt, 1.0, None, False, None, None, None, None, dw2, False, None, 0, 1)
    torch.testing.assert_close(dw2, dz.float().t() @ a.float(), rtol=2e-3, atol=5e-2)


def test_gemm_fused_ce_epilogue(C):
    # logits = a W^T + b, then softmax-CE in the SAME kernel: loss, dZ (bf16, zero padded), db
    Bt, I, Cn = 300, 256, 10
    a, w = _mk(Bt, I, 0.2, 40), _mk(Cn, I, 0.2, 41)
    b = torch.randn(Cn, device=dev())
    y = torch.randint(0, Cn, (Bt,), device=dev(), dtype=torch.int32)
    loss = torch.zeros(1, device=dev())
    dz = torch.full((Bt, 16), 7.0, dtype=torch.bfloat16, device=dev())[:, :Cn]
    db = torch.zeros(Cn, device=dev())
    logits = torch.zeros(Bt, 16, device=dev())[:, :Cn]
    C.gemm_bf16(a, False, w, False, Bt, Cn, I, bias=b, out_bf16=dz, colsum=db, out_f32=logits, loss_kind=1, labels=y,
                loss_slot=loss, loss_scale=1.0 / Bt)
    zr = (a.float() @ w.float().t() + b).requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(zr, y.long())
    ref.backward()
    torch.testing.assert_close(logits, zr.detach(), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dz.float(), zr.grad, rtol=2e-2, atol=1e-5)
    full = dz.as_strided((Bt, 16), (16, 1))
    assert full[:, Cn:].abs().max().item() == 0
    torch.testing.assert_close(db, zr.grad.sum(0), rtol=1e-2, atol=1e-4)
    # loss-only (validation) mode
    l2 = torch.zeros(1, device=dev())
    C.gemm_bf16(a, False, w, False, Bt, Cn, I, bias=b, loss_kind=1, labels=y, loss_slot=l2, loss_scale=1.0 / Bt)
    torch.testing.assert_close(l2[0], ref.detach(), rtol=1e-4, atol=1e-4)


def test_gemm_fused_mse_epilogue(C):
    Bt, I, D = 200, 20, 3
    a, w = _mk(Bt, I, 0.5, 42), _mk(D, I, 0.5, 43)
    t = torch.randn(Bt, D, device=dev())
    loss = torch.zeros(1, device=dev())
    dz = torch.zeros(Bt, 8, dtype=torch.bfloat16, device=dev())[:, :D]
    db = torch.zeros(D, device=dev())
    C.gemm_bf16(a, False, w, False, Bt, D, I, out_bf16=dz, colsum=db, loss_kind=2, target=t, loss_slot=loss,
                loss_scale=1.0 / (Bt * D))
    pr = (a.float() @ w.float().t()).requires_grad_(True)
    ref = torch.nn.functional.mse_loss(pr, t)
    ref.backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dz.float(), pr.grad, rtol=2e-2, atol=1e-5)
    torch.testing.assert_close(db, pr.grad.sum(0), rtol=1e-2, atol=1e-4)


def test_gemm_device_side_batch_windows(C):
    # the batch is window (iter % n_windows) of a resident partition, selected inside the kernel
    n, Bt, I, O, Cn = 1024, 256, 784, 256, 10
    X = _mk(n, I, 0.1, 44)
    w = _mk(O, I, 0.1, 45)
    state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev())
    for it in (0, 3, 6):
        state.view(torch.int64)[2] = it
        off = (it % (n // Bt)) * Bt
        y = torch.zeros(Bt, O, dtype=torch.bfloat16, device=dev())
        C.gemm_bf16(X, False, w, False, Bt, O, I, out_bf16=y, win_state=state, window_rows=Bt, n_windows=n // Bt, win_flags=1)
        torch.testing.assert_close(y.float(), X[off:off + Bt].float() @ w.float().t(), rtol=2e-2, atol=2e-2)
        # wgrad with the windowed partition as MN-major B operand: dW = dz^T X[window]
        dzz = _mk(Bt, O, 0.05, 46)
        dw = torch.zeros(O, I, device=dev())
        C.gemm_bf16(dzz, True, X, True, O, I, Bt, out_f32=dw, win_state=state, window_rows=Bt, n_windows=n // Bt, win_flags=2)
        torch.testing.assert_close(dw, dzz.float().t() @ X[off:off + Bt].float(), rtol=2e-3, atol=2e-2)
        # windowed labels for the fused loss
        labels = torch.randint(0, Cn, (n,), device=dev(), dtype=torch.int32)
        a, wc = _mk(Bt, O, 0.2, 47), _mk(Cn, O, 0.2, 48)
        loss = torch.zeros(1, device=dev())
        C.gemm_bf16(a, False, wc, False, Bt, Cn, O, loss_kind=1, labels=labels, loss_slot=loss, loss_scale=1.0 / Bt,
                    win_state=state, window_rows=Bt, n_windows=n // Bt, win_flags=4)
        ref = torch.nn.functional.cross_entropy(a.float() @ wc.float().t(), labels[off:off + Bt].long())
        torch.testing.assert_close(loss[0], ref, rtol=1e-4, atol=1e-4)


def test_softmax_ce_kernel(C):
    Bt, Cn = 300, 10
    z = torch.randn(Bt, 16, device=dev())[:, :Cn] * 3
    y = torch.randint(0, Cn, (Bt,), device=dev(), dtype=torch.int32)
    loss = torch.zeros(1, device=dev())
    dz = torch.full((Bt, 16), 7.0, dtype=torch.bfloat16, device=dev())
    dzt = torch.zeros(Cn, _pad8(Bt), dtype=torch.bfloat16, device=dev())[:, :Bt]
    db = torch.zeros(Cn, device=dev())
    C.softmax_ce(z, y, Cn, loss, dz, dzt, db, 16)
    zr = z.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(zr, y.long())
    ref.backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dz[:, :Cn].float(), zr.grad, rtol=1e-2, atol=1e-5)
    assert dz[:, Cn:].abs().max().item() == 0
    torch.testing.assert_close(dzt.float(), zr.grad.t(), rtol=1e-2, atol=1e-5)
    torch.testing.assert_close(db, zr.grad.sum(0), rtol=1e-3, atol=1e-5)
    # loss-only mode (validation)
    l2 = torch.zeros(1, device=dev())
    C.softmax_ce(z, y, Cn, l2, None, None, None, 16)
    torch.testing.assert_close(l2[0], ref.detach(), rtol=1e-5, atol=1e-5)


def test_loss_kernels_read_targets_of_the_current_window(C):
    """'sweep' batches: labels / targets stay in the resident partition, the kernel offsets them by the device-side
    iteration counter exactly like the GEMM batch windows."""
    from sparktorch_b200.ops import reference

    Bt, Cn, windows, it = 8, 10, 3, 5
    state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev())
    state[4] = it                                               # StepState.iter
    off = (it % windows) * Bt
    z = torch.randn(Bt, 16, device=dev())[:, :Cn]
    y_all = torch.randint(0, Cn, (Bt * windows,), device=dev(), dtype=torch.int32)
    loss = torch.zeros(1, device=dev())
    dz = torch.zeros(Bt, 16, dtype=torch.bfloat16, device=dev())
    C.softmax_ce(z, y_all, Cn, loss, dz, None, None, 16, win_state=state, window_rows=Bt, n_windows=windows)
    ref = torch.nn.functional.cross_entropy(z, y_all[off: off + Bt].long())
    torch.testing.assert_close(loss[0], ref, rtol=1e-5, atol=1e-5)
    l_ref, dz_ref = torch.zeros(1), torch.zeros(Bt, 16)
    reference.softmax_ce(z.cpu(), y_all.cpu(), Cn, l_ref, dz_ref, None, None, 16, win_state=state.cpu(), window_rows=Bt, n_windows=windows)
    torch.testing.assert_close(dz.float().cpu(), dz_ref, rtol=1e-2, atol=1e-5)
    D = 3
    p = torch.randn(Bt, 8, device=dev())[:, :D]
    t_all = torch.randn(Bt * windows, D, device=dev())
    l2 = torch.zeros(1, device=dev())
    C.mse_loss(p, t_all, D, l2, None, None, None, 8, win_state=state, window_rows=Bt, n_windows=windows)
    torch.testing.assert_close(l2[0], torch.nn.functional.mse_loss(p, t_all[off: off + Bt]), rtol=1e-5, atol=1e-6)


def test_mse_kernel(C):
    Bt, D = 200, 3
    p = torch.randn(Bt, 8, device=dev())[:, :D]
    t = torch.randn(Bt, D, device=dev())
    loss = torch.zeros(1, device=dev())
    dz = torch.zeros(Bt, 8, dtype=torch.bfloat16, device=dev())
    db = torch.zeros(D, device=dev())
    C.mse_loss(p, t, D, loss, dz, None, db, 8)
    pr = p.clone().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(pr, t)
    ref.backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dz[:, :D].float(), pr.grad, rtol=1e-2, atol=1e-6)
    torch.testing.assert_close(db, pr.grad.sum(0), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gather_batch(C, mode):
    n, D, Bt = 1000, 784, 256
    src = _mk(n, D, 1.0, 20)
    labels = torch.randint(0, 10, (n,), device=dev(), dtype=torch.int32)
    state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev())
    it = 3
    state.view(torch.int64)[2] = it      # StepState.iter (bytes 16..23)
    x = torch.zeros(Bt, _pad8(D), dtype=torch.bfloat16, device=dev())[:, :D]
    xt = torch.zeros(D, Bt, dtype=torch.bfloat16, device=dev())
    yo = torch.zeros(Bt, dtype=torch.int32, device=dev())
    ring = 3
    C.gather_batch(src, n, labels, 0, mode, ring, 1234, Bt, D, x, xt, yo, state)
    torch.cuda.synchronize()
    # recover the row choice by matching labels + content
    if mode == 0:
        idx = (torch.arange(Bt, device=dev()) + it * Bt) % n
    elif mode == 2:
        idx = torch.arange(Bt, device=dev()) + (it % ring) * Bt
    else:
        # find each gathered row in src (rows are unique random vectors)
        sim = (x.float() @ src.float().t())
        idx = sim.argmax(1)
        assert idx.unique().numel() == Bt, "sampling must be without replacement"
    torch.testing.assert_close(x.float(), src[idx].float(), rtol=0, atol=0)
    torch.testing.assert_close(xt.float(), src[idx].float().t(), rtol=0, atol=0)
    assert torch.equal(yo, labels[idx])


def test_feistel_sampling_changes_per_iteration_and_covers(C):
    n, D, Bt = 512, 16, 128
    src = torch.zeros(n, D, dtype=torch.bfloat16, device=dev())
    src[:, 0] = torch.arange(n, device=dev()).to(torch.bfloat16)   # exact up to 256; use two columns
    src[:, 1] = (torch.arange(n, device=dev()) // 256).to(torch.bfloat16)
    src[:, 0] = (torch.arange(n, device=dev()) % 256).to(torch.bfloat16)
    state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev())
    x = torch.zeros(Bt, D, dtype=torch.bfloat16, device=dev())
    seen = []
    for it in range(4):
        state.view(torch.int64)[2] = it
        C.gather_batch(src, n, None, 0, 1, 1, 99, Bt, D, x, None, None, state)
        ids = (x[:, 0].float() + 256 * x[:, 1].float()).long()
        assert ids.unique().numel() == Bt
        seen.append(set(ids.tolist()))
    assert seen[0] != seen[1] and seen[1] != seen[2]


def test_transpose_kernel(C):
    src = _mk(70, 130, 1.0, 30)
    dst = torch.zeros(130, 72, dtype=torch.bfloat16, device=dev())
    C.transpose_bf16(src, 70, 130, dst)
    torch.testing.assert_close(dst[:, :70].float(), src.float().t(), rtol=0, atol=0)


def _updater(C, spec, numel, grad, param, wb, segs_rows, patience=-1, zero=True, iters=64):
    from sparktorch_b200.runtime.flat import TAIL

    state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev())
    m = torch.zeros(numel, device=dev()) if spec.n_state >= 1 else None
    v = torch.zeros(numel, device=dev()) if spec.n_state >= 2 else None
    hist = torch.zeros(iters, device=dev())
    segs = torch.frombuffer(bytearray(C.pack_segments(segs_rows)), dtype=torch.uint8).to(dev())
    mail = torch.zeros(4, dtype=torch.int32).pin_memory()
    up = C.Updater(1, 0, [grad.data_ptr()], [param.data_ptr()], [wb.data_ptr()] if wb is not None else [], [], 0, m, v,
                   state, hist, mail.data_ptr(), segs, numel, TAIL, spec.kind_id, spec.lr, spec.beta1, spec.beta2, spec.eps,
                   spec.weight_decay, spec.momentum, spec.dampening, spec.nesterov, spec.maximize, patience, zero, 0, 0, 0, 0)
    return up, state, hist, mail, (m, v, segs)


OPT_CASES = [
    ("adam", dict(lr=1e-2)),
    ("adam", dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)),
    ("adamw", dict(lr=1e-2, weight_decay=0.05)),
    ("sgd", dict(lr=0.1)),
    ("sgd", dict(lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.01)),
    ("sgd", dict(lr=0.1, momentum=0.5, dampening=0.3)),
]


@pytest.mark.parametrize("name,kw", OPT_CASES)
def test_fused_update_matches_torch_optim(C, name, kw):
    from sparktorch_b200.runtime.flat import TAIL
    from sparktorch_b200.runtime.optim import parse_optimizer

    cls = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD}[name]
    spec = parse_optimizer(cls, kw)
    rows, cols = 20, 10
    numel = 256                                    # 200 weight (padded to 256)
    torch.manual_seed(0)
    p_ref = torch.nn.Parameter(torch.randn(rows, cols, device=dev()))
    opt = cls([p_ref], **kw)
    param = torch.zeros(numel, device=dev())
    param[:200] = p_ref.detach().reshape(-1)
    grad = torch.zeros(numel + TAIL, device=dev())
    wb = torch.zeros(rows * 16 + 16 * 24, dtype=torch.bfloat16, device=dev())
    up, state, hist, mail, keep = _updater(C, spec, numel, grad, param, wb, [[0, rows, cols, 0, 16, rows * 16, 24, 1]])
    for t in range(6):
        g = torch.randn(rows, cols, device=dev())
        grad[:200] = g.reshape(-1)
        grad[numel] = float(t)                      # loss slot
        p_ref.grad = g.clone()
        opt.step()
        up.step()
        torch.cuda.synchronize()
        torch.testing.assert_close(param[:200].view(rows, cols), p_ref.detach(), rtol=3e-5, atol=3e-6)
        assert grad.abs().max().item() == 0        # arena cleared for the next step
    w = wb[: rows * 16].view(rows, 16)[:, :cols]
    wt = wb[rows * 16:].view(16, 24)[:cols, :rows]
    torch.testing.assert_close(w.float(), p_ref.detach().to(torch.bfloat16).float(), rtol=0, atol=0)
    torch.testing.assert_close(wt.float(), p_ref.detach().t().to(torch.bfloat16).float(), rtol=0, atol=0)
    torch.testing.assert_close(hist[:6].cpu(), torch.arange(6.0))
    assert int(mail[1]) == 6 and int(mail[0]) == 0
    assert int(state.view(torch.int64)[1]) == 6    # StepState.step


def test_fused_update_early_stop_in_band(C):
    from sparktorch_b200.runtime.flat import TAIL
    from sparktorch_b200.runtime.optim import parse_optimizer

    spec = parse_optimizer(torch.optim.SGD, dict(lr=0.5))
    numel = 64
    param = torch.ones(numel, device=dev())
    grad = torch.zeros(numel + TAIL, device=dev())
    up, state, hist, mail, keep = _updater(C, spec, numel, grad, param, None, [[0, 1, 64, -1, 0, -1, 0, 1]], patience=2)
    losses = [1.0, 0.5, 0.6, 0.7, 0.1, 0.1]          # seeds, improves, bad, bad -> stop at iteration 3
    applied = []
    for t, l in enumerate(losses):
        grad[:numel] = 1.0
        grad[numel] = l
        up.step()
        torch.cuda.synchronize()
        applied.append(float(param[0]))
    # updates at t=0,1,2 ; t=3 triggers the stop and is NOT applied (reference semantics) ; later launches are no-ops
    assert applied == [0.5, 0.0, -0.5, -0.5, -0.5, -0.5]
    assert int(mail[0]) == 1 and int(mail[1]) == 4
    assert int(state.view(torch.int64)[1]) == 3


def test_publish_only_builds_bf16_copies(C):
    from sparktorch_b200.runtime.flat import TAIL
    from sparktorch_b200.runtime.optim import parse_optimizer

    spec = parse_optimizer(torch.optim.Adam, {})
    rows, cols, numel = 10, 256, 2560
    param = torch.randn(numel, device=dev())
    grad = torch.zeros(numel + TAIL, device=dev())
    wb = torch.zeros(rows * cols, dtype=torch.bfloat16, device=dev())
    up, *_ = _updater(C, spec, numel, grad, param, wb, [[0, rows, cols, 0, cols, -1, 0, 0]])
    before = param.clone()
    up.publish()
    torch.cuda.synchronize()
    assert torch.equal(param, before)
    torch.testing.assert_close(wb.view(rows, cols).float(), param.view(rows, cols).to(torch.bfloat16).float(), rtol=0, atol=0)


def test_im2col_matches_unfold(C):
    import torch.nn.functional as F

    B, Cin, H, W, kh, kw = 5, 3, 12, 10, 3, 3
    for stride, pad in ((1, 0), (1, 1), (2, 1), (1, 2)):
        x = torch.randn(B, Cin, H, W, device=dev()).to(torch.bfloat16)
        ref = F.unfold(x.float(), (kh, kw), padding=pad, stride=stride)          # [B, Cin*kh*kw, L]
        OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        K = Cin * kh * kw
        # NCHW source (torch row layout)
        col = torch.full((B * OH * OW, _pad8(K)), 9.0, dtype=torch.bfloat16, device=dev())[:, :K]
        xf = x.reshape(B, -1)
        C.im2col_nhwc(xf, B, H, W, Cin, xf.stride(0), W, 1, H * W, kh, kw, stride, pad, col)
        got = col.float().view(B, OH * OW, K).permute(0, 2, 1)
        torch.testing.assert_close(got, ref, rtol=0, atol=0)
        # NHWC source
        xn = x.permute(0, 2, 3, 1).contiguous()
        col2 = torch.zeros_like(col)
        C.im2col_nhwc(xn, B, H, W, Cin, H * W * Cin, W * Cin, Cin, 1, kh, kw, stride, pad, col2)
        torch.testing.assert_close(col2.float(), col.float(), rtol=0, atol=0)


def test_conv_forward_and_backward_via_gemm(C):
    """conv fwd = im2col+GEMM; wgrad = dY^T col; dgrad = im2col(pad) of dY x flipped taps  vs  autograd."""
    import torch.nn.functional as F

    B, Cin, H, W, Cout, k, pad = 4, 16, 12, 12, 32, 3, 1
    x = (torch.randn(B, Cin, H, W, device=dev()) * 0.5).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, device=dev()) * 0.1).to(torch.bfloat16)
    b = torch.randn(Cout, device=dev())
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, b, padding=pad)
    gy = (torch.randn_like(y_ref) * 0.1).to(torch.bfloat16)
    y_ref.backward(gy.float())
    OH = OW = H + 2 * pad - k + 1
    K = Cin * k * k
    xn = x.permute(0, 2, 3, 1).contiguous()                       # NHWC
    col = torch.zeros(B * OH * OW, K, dtype=torch.bfloat16, device=dev())
    C.im2col_nhwc(xn, B, H, W, Cin, H * W * Cin, W * Cin, Cin, 1, k, k, 1, pad, col)
    w2 = w.reshape(Cout, K).contiguous()
    y = torch.zeros(B * OH * OW, Cout, dtype=torch.bfloat16, device=dev())
    C.gemm_bf16(col, False, w2, False, B * OH * OW, Cout, K, bias=b, out_bf16=y)
    torch.testing.assert_close(y.float().view(B, OH, OW, Cout).permute(0, 3, 1, 2), y_ref.detach(), rtol=3e-2, atol=3e-2)
    # wgrad
    gyn = gy.permute(0, 2, 3, 1).reshape(B * OH * OW, Cout).contiguous()
    dw = torch.zeros(Cout, K, device=dev())
    C.gemm_bf16(gyn, True, col, True, Cout, K, B * OH * OW, out_f32=dw, atomic_f32=True, split_k=4)
    torch.testing.assert_close(dw.view(Cout, Cin, k, k), wr.grad, rtol=3e-2, atol=3e-2)
    # dgrad
    wf = torch.zeros(Cin, Cout * k * k, dtype=torch.bfloat16, device=dev())
    C.flip_conv_weight(w2, Cout, Cin, k, k, wf)
    colT = torch.zeros(B * H * W, Cout * k * k, dtype=torch.bfloat16, device=dev())
    C.im2col_nhwc(gyn, B, OH, OW, Cout, OH * OW * Cout, OW * Cout, Cout, 1, k, k, 1, k - 1 - pad, colT)
    dx = torch.zeros(B * H * W, Cin, dtype=torch.bfloat16, device=dev())
    C.gemm_bf16(colT, False, wf, False, B * H * W, Cin, Cout * k * k, out_bf16=dx)
    torch.testing.assert_close(dx.float().view(B, H, W, Cin).permute(0, 3, 1, 2), xr.grad, rtol=3e-2, atol=3e-2)


def test_maxpool_dropout_kernels(C):
    import torch.nn.functional as F

    B, Cc, H, W, k = 6, 32, 22, 22, 2
    x = torch.randn(B, Cc, H, W, device=dev()).to(torch.bfloat16)
    xn = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, Cc)
    PH, PW = H // k, W // k
    y = torch.zeros(B, Cc * PH * PW, dtype=torch.bfloat16, device=dev())
    idx = torch.zeros(B * PH * PW * Cc, dtype=torch.uint8, device=dev())
    C.maxpool_dropout(False, xn, Cc, y, idx, B, H, W, Cc, k)
    xr = x.float().requires_grad_(True)
    ref = F.max_pool2d(xr, k)
    torch.testing.assert_close(y.float().view(B, Cc, PH, PW), ref.detach(), rtol=0, atol=0)
    gy = torch.randn(B, Cc * PH * PW, device=dev()).to(torch.bfloat16)
    ref.backward(gy.float().view(B, Cc, PH, PW))
    dx = torch.full((B * H * W, Cc), 5.0, dtype=torch.bfloat16, device=dev())
    db = torch.zeros(Cc, device=dev())
    C.maxpool_dropout(True, dx, Cc, gy, idx, B, H, W, Cc, k, db=db)
    torch.testing.assert_close(dx.float().view(B, H, W, Cc).permute(0, 3, 1, 2), xr.grad, rtol=0, atol=0)
    torch.testing.assert_close(db, xr.grad.sum((0, 2, 3)), rtol=1e-3, atol=1e-3)
    # NHWC pooled layout + dropout2d statistics
    y2 = torch.zeros(B * PH * PW, Cc, dtype=torch.bfloat16, device=dev())
    C.maxpool_dropout(False, xn, Cc, y2, idx, B, H, W, Cc, k, p_drop=0.5, train=True, seed=7, y_nhwc=True)
    kept = (y2.float().view(B, PH * PW, Cc).abs().sum(1) > 0)
    frac = kept.float().mean().item()
    assert 0.3 < frac < 0.7
    ref2 = ref.detach().permute(0, 2, 3, 1).reshape(B, PH * PW, Cc) * 2.0
    torch.testing.assert_close(y2.float().view(B, PH * PW, Cc)[kept.unsqueeze(1).expand(-1, PH * PW, -1)],
                               ref2[kept.unsqueeze(1).expand(-1, PH * PW, -1)], rtol=1e-2, atol=1e-2)


def test_gemm_k_ag_gathers_weights_from_sharded_fp32_arena(C):
    """K-AG on one GPU: the flat fp32 arena is 'owned' in 4 shards (4 separate buffers standing in for 4 GPUs);
    the GEMM gathers + casts the weight tile itself, adds the gathered bias and fills the local bf16 copy."""
    world, Bt, I, O = 4, 256, 784, 256
    numel = ((O * I + 63) // 64) * 64 + 256
    shard = ((numel // 4 + world - 1) // world) * 4
    flat = torch.randn(numel, device=dev()) * 0.05
    # every "owner" holds a full-size buffer but only its shard is valid (others poisoned)
    owners = []
    for r in range(world):
        b = torch.full((numel,), float("nan"), device=dev())
        b[r * shard:(r + 1) * shard] = flat[r * shard:(r + 1) * shard]
        owners.append(b)
    w_off, b_off = 0, ((O * I + 63) // 64) * 64
    W = flat[w_off:w_off + O * I].view(O, I)
    bias = flat[b_off:b_off + O]
    x = _mk(Bt, I, 0.2, 60)
    y = torch.zeros(Bt, O, dtype=torch.bfloat16, device=dev())
    wb = torch.zeros(O, I, dtype=torch.bfloat16, device=dev())
    C.gemm_bf16(x, False, wb, False, Bt, O, I, relu=True, out_bf16=y, peer_table=torch.tensor([o.data_ptr() for o in owners], dtype=torch.int64, device=dev()),
                peer_shard_elems=shard, peer_w_off=w_off, peer_bias_off=b_off, peer_fill_local=True)
    ref = torch.relu(x.float() @ W.to(torch.bfloat16).float().t() + bias)
    torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(wb.float(), W.to(torch.bfloat16).float(), rtol=0, atol=0)
    # small head with ragged K and fused CE
    O2, I2 = 10, 250
    flat2 = torch.randn(4096, device=dev()) * 0.1
    sh2 = 1024
    owners2 = [flat2.clone() for _ in range(4)]
    a = _mk(Bt, I2, 0.3, 61)
    labels = torch.randint(0, O2, (Bt,), device=dev(), dtype=torch.int32)
    loss = torch.zeros(1, device=dev())
    wb2 = torch.zeros(O2, 256, dtype=torch.bfloat16, device=dev())[:, :I2]
    C.gemm_bf16(a, False, wb2, False, Bt, O2, I2, loss_kind=1, labels=labels, loss_slot=loss, loss_scale=1.0 / Bt,
                peer_table=torch.tensor([o.data_ptr() for o in owners2], dtype=torch.int64, device=dev()), peer_shard_elems=sh2, peer_w_off=64, peer_bias_off=3008)
    W2 = flat2[64:64 + O2 * I2].view(O2, I2).to(torch.bfloat16).float()
    ref2 = torch.nn.functional.cross_entropy(a.float() @ W2.t() + flat2[3008:3008 + O2], labels.long())
    torch.testing.assert_close(loss[0], ref2, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_persistent_variant(C, a_mn, b_mn, block_n):
    """Persistent CTAs + double-buffered TMEM accumulator: many tiles per CTA, ragged edges, fused epilogue."""
    M, N, K = 2000, 1100, 520
    A = _mk(K, M, 0.5, 70) if a_mn else _mk(M, K, 0.5, 70)
    B = _mk(K, N, 0.5, 71) if b_mn else _mk(N, K, 0.5, 71)
    bias = torch.randn(N, device=dev())
    mask = _mk(M, N, 1.0, 72)
    mask.clamp_(min=0)                       # keep the padded pitch (torch.relu would return a dense copy)
    ref = ((A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t()) + bias)
    ref = torch.relu(ref) * (mask.float() > 0)
    out = torch.zeros(M, _pad8(N), dtype=torch.bfloat16, device=dev())[:, :N]
    o32 = torch.zeros(M, _pad8(N), dtype=torch.float32, device=dev())[:, :N]
    cs = torch.zeros(N, device=dev())
    C.set_persistent_mode(2)
    try:
        C.gemm_bf16(A, a_mn, B, b_mn, M, N, K, bias=bias, relu=True, mask_src=mask, out_bf16=out, out_f32=o32, colsum=cs,
                    block_n=block_n)
        torch.cuda.synchronize()
    finally:
        C.set_persistent_mode(1)
    torch.testing.assert_close(o32, ref, rtol=2e-3, atol=0.15)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=0.2)
    torch.testing.assert_close(cs, ref.sum(0), rtol=2e-2, atol=2.0)


@pytest.mark.parametrize("M,N,K,bias,relu", [(256, 256, 64, False, False), (1000, 640, 200, True, True), (2048, 1024, 4096 + 64, True, False),
                                             (777, 1280, 512, False, True)])
def test_gemm_cta_pair_variant(C, M, N, K, bias, relu):
    """tcgen05.mma.cta_group::2 kernel (CTA pairs, 256 x 256 tiles, csrc/gemm2cta_sm100.cu): ragged M / N / K, ring wrap-around
    (K = 65 k-blocks > 6 stages), both accumulator stages (more tiles than pairs is covered by 2048 x 1024 on a full GPU)."""
    a, b = _mk(M, K, seed=1), _mk(N, K, seed=2)
    bv = torch.randn(N, device=dev()) if bias else None
    out = torch.full((M, _pad8(N)), float("nan"), dtype=torch.bfloat16, device=dev())[:, :N]
    rc = C.gemm_bf16_2cta(a, b, out, bv, 0.5, relu)
    torch.cuda.synchronize()
    assert rc == 0
    ref = 0.5 * (a.float() @ b.float().t())
    if bias:
        ref = ref + bv
    if relu:
        ref = ref.clamp_min(0)
    assert torch.isfinite(out.float()).all()
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2 * math.sqrt(K) / 8 + 1e-2)


def test_gemm_cta_pair_declines_small_shapes(C):
    a, b = _mk(128, 64), _mk(64, 64)
    out = torch.zeros(128, 64, dtype=torch.bfloat16, device=dev())
    assert C.gemm_bf16_2cta(a, b, out, None, 1.0, False) == -2


@pytest.mark.skipif(__import__("os").environ.get("SPARKTORCH_IMPLICIT_CONV", "0") != "1",
                    reason="implicit-GEMM convolution (TMA im2col) was written after the round's GPU budget was spent: run "
                           "bench/conv_implicit_check.py first, then enable with SPARKTORCH_IMPLICIT_CONV=1")
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad", [(2, 8, 8, 64, 64, 3, 1, 1), (4, 14, 14, 128, 64, 3, 1, 1), (2, 16, 16, 64, 128, 3, 2, 1),
                                                         (2, 12, 12, 64, 64, 1, 1, 0), (3, 10, 7, 64, 32, 3, 1, 1)])
def test_conv_forward_implicit_gemm(C, N, H, W, Cin, Cout, k, stride, pad):
    import torch.nn.functional as F

    x = (torch.randn(N, Cin, H, W, device=dev()) * 0.5).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, device=dev()) * 0.05).to(torch.bfloat16)
    b = torch.randn(Cout, device=dev())
    P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
    out = torch.full((N * P * Q, Cout), float("nan"), dtype=torch.bfloat16, device=dev())
    assert C.conv2d_fwd_implicit(xn, wk, out, b, True, k, k, stride, pad, 1) == 0
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float(), w.float(), b, stride=stride, padding=pad)).permute(0, 2, 3, 1).reshape(N * P * Q, Cout)
    torch.testing.assert_close(out.float(), ref, rtol=3e-2, atol=3e-2)
