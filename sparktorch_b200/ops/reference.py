"""Plain-PyTorch reference implementations of the native ops (same names / arguments as ``_C``).

Two uses:

* an **oracle** for the kernel tests (every sm_100a kernel is compared against fp32 PyTorch of the same op);
* an **emulation back-end** for the fused plans: ``FusedResNet`` & co. only talk to the extension through a
  handful of calls (``gemm_bf16``, ``im2col_cl``, ``bn_apply`` ...), so swapping this module in runs the *same
  orchestration code* (buffer wiring, gradient routing through residual joins, arena offsets) on the CPU in fp32,
  where it can be checked exactly against autograd (tests/test_plan_emulation.py) without a GPU.

Outputs are written in place into the caller's (possibly padded / strided) buffers, like the kernels do.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32)


# ----------------------------------------------------------------------------- GEMM
def _window_offset(win_state, window_rows, n_windows) -> int:
    it = int(win_state[4]) + (int(win_state[5]) << 32)            # StepState.iter (int64 at int32 words 4..5)
    return (it % max(1, n_windows)) * window_rows


def gemm_bf16(A, a_mn, B, b_mn, M, N, K, alpha: float = 1.0, bias=None, relu: bool = False, mask_src=None, out_bf16=None,
              out_t_bf16=None, out_f32=None, out_t_f32=None, atomic_f32: bool = False, colsum=None, block_n: int = 0,
              split_k: int = 1, loss_kind: int = 0, labels=None, target=None, loss_slot=None, loss_scale: float = 1.0,
              win_state=None, window_rows: int = 0, n_windows: int = 0, win_flags: int = 0, colsumsq=None, peer_table=None,
              peer_shard_elems: int = 0, peer_w_off: int = 0, peer_bias_off: int = -1, peer_fill_local: bool = False, **unsupported):
    """D[M, N] = A · Bᵀ with the epilogue options the plans use (A stored [M,K], or [K,M] when ``a_mn``; same for B).

    ``win_flags``: 1 - A is the resident partition, the batch is the window of the current iteration; 2 - same for the
    reduction rows of an MN-major B; 4 - labels / targets are taken from the same window.
    ``loss_kind`` 1 (softmax-CE, mean over rows) / 2 (MSE, mean over M*N): the accumulator becomes dLoss/dOutput, so
    ``out_bf16`` / ``colsum`` emit dZ / db while ``out_f32`` keeps the raw output.
    ``peer_table`` (K-AG, here a list of per-owner fp32 arenas): B[n, k] and the bias are gathered from the owner of flat
    element ``f = peer_w_off + n*K + k`` (owner = f // peer_shard_elems); ``peer_fill_local`` leaves the gathered
    weights in ``B`` for the dgrad GEMM."""
    for k, v in unsupported.items():
        if v not in (None, 0, False, 0.0, 1.0):
            raise NotImplementedError(f"reference gemm_bf16: option {k} is not emulated")
    off = _window_offset(win_state, window_rows, n_windows) if (win_state is not None and win_flags) else 0
    a_off = off if (win_flags & 1) else 0
    b_off = off if (win_flags & 2) else 0
    l_off = off if (win_flags & 4) else 0
    Am = _f(A[:K, :M]).t() if a_mn else _f(A[a_off: a_off + M, :K])
    if peer_table is not None:
        assert not b_mn
        f = peer_w_off + torch.arange(N * K)
        owner = f // peer_shard_elems
        flat = torch.stack([_f(t_) for t_ in peer_table])            # [owners, numel]
        Bm = flat[owner, f].reshape(N, K)
        if peer_fill_local:
            B[:N, :K] = Bm.to(B.dtype)
        if peer_bias_off >= 0:
            fb = peer_bias_off + torch.arange(N)
            bias = flat[fb // peer_shard_elems, fb]
    else:
        Bm = _f(B[b_off: b_off + K, :N]).t() if b_mn else _f(B[:N, :K])
    D = (Am @ Bm.t()) * alpha
    if bias is not None:
        D = D + _f(bias[:N])
    if relu:
        D = torch.relu(D)
    if mask_src is not None:
        D = D * (_f(mask_src[:M, :N]) > 0)
    if loss_kind:
        if out_f32 is not None:
            out_f32[:M, :N] = D
        if loss_kind == 1:
            lab = labels[l_off: l_off + M].to(torch.int64)
            loss_slot[0] += F.cross_entropy(D, lab, reduction="sum") * loss_scale
            D = (torch.softmax(D, 1) - F.one_hot(lab, N)) * loss_scale
        else:
            dlt = D - _f(target[l_off: l_off + M, :N])
            loss_slot[0] += (dlt * dlt).sum() * loss_scale
            D = 2.0 * dlt * loss_scale
    if colsum is not None:
        colsum[:N] += D.sum(0)
    if colsumsq is not None:
        colsumsq[:N] += (D * D).sum(0)
    if out_bf16 is not None:
        out_bf16[:M, :N] = D.to(out_bf16.dtype)
    if out_t_bf16 is not None:
        out_t_bf16[:N, :M] = D.t().to(out_t_bf16.dtype)
    if out_f32 is not None and not loss_kind:
        if atomic_f32:
            out_f32[:M, :N] += D
        else:
            out_f32[:M, :N] = D
    if out_t_f32 is not None:
        if atomic_f32:
            out_t_f32[:N, :M] += D.t()
        else:
            out_t_f32[:N, :M] = D.t()


# ----------------------------------------------------------------------------- im2col variants
def _unfold_to(x_nchw: torch.Tensor, k: int, stride: int, pad: int, OH: int, OW: int) -> torch.Tensor:
    """[B, C, H, W] -> [B, C, k*k, OH*OW] patches with zero padding, for an explicitly requested output size."""
    B, C, H, W = x_nchw.shape
    need_h, need_w = (OH - 1) * stride + k, (OW - 1) * stride + k
    xp = F.pad(x_nchw, (pad, max(0, need_w - W - pad), pad, max(0, need_h - H - pad)))[:, :, :need_h, :need_w]
    cols = F.unfold(xp, k, stride=stride)                     # [B, C*k*k, OH*OW] with (c, i, j) rows
    return cols.reshape(B, C, k * k, OH * OW)


def im2col_nhwc(x, B, H, W, C, sb, sh, sw, sc, kh, kw, stride, pad, col, state=None, window_rows: int = 0, n_windows: int = 0):
    """Generic-stride source, (c, i, j) column order (= torch's [Cout, Cin, kh, kw] weight layout); with ``state`` the
    samples are the current iteration's window of a resident partition."""
    assert kh == kw
    b_off = _window_offset(state, window_rows, n_windows) if state is not None else 0
    src = torch.as_strided(x, (B, C, H, W), (sb, sc, sh, sw), x.storage_offset() + b_off * sb)
    OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    p = _unfold_to(_f(src), kh, stride, pad, OH, OW)           # [B, C, kk, L]
    col[: B * OH * OW, : C * kh * kw] = p.permute(0, 3, 1, 2).reshape(B * OH * OW, C * kh * kw).to(col.dtype)


def im2col_cl(x, B, H, W, C, kh, kw, stride, pad, dil, OH, OW, col):
    """Channels-last source [B*H*W, C], (i, j, c) column order, optional zero insertion (``dil``) for strided dgrad."""
    assert kh == kw
    xn = _f(x[: B * H * W, :C]).reshape(B, H, W, C).permute(0, 3, 1, 2)
    if dil > 1:
        xd = xn.new_zeros(B, C, (H - 1) * dil + 1, (W - 1) * dil + 1)
        xd[:, :, ::dil, ::dil] = xn
        xn = xd
    p = _unfold_to(xn, kh, stride, pad, OH, OW)                # [B, C, kk, L]
    col[: B * OH * OW, : kh * kw * C] = p.permute(0, 3, 2, 1).reshape(B * OH * OW, kh * kw * C).to(col.dtype)


def flip_conv_weight(w, Cout, Cin, kh, kw, out):
    """[Cout, (ci, i, j)] -> [Cin, (co, kh-1-i, kw-1-j)]"""
    w4 = w[:Cout, : Cin * kh * kw].reshape(Cout, Cin, kh, kw)
    out[:Cin, : Cout * kh * kw] = w4.flip(2, 3).permute(1, 0, 2, 3).reshape(Cin, Cout * kh * kw)


def flip_conv_weight_cl(w, Cout, Cin, kh, kw, out):
    w4 = w[:Cout, : kh * kw * Cin].reshape(Cout, kh * kw, Cin)
    out[:Cin, : kh * kw * Cout] = w4.flip(1).permute(2, 1, 0).reshape(Cin, kh * kw * Cout)


# ----------------------------------------------------------------------------- batch norm
def bn_stats(y, M, C, sums):
    yf = _f(y[:M, :C])
    sums[:C] += yf.sum(0)
    sums[C: 2 * C] += (yf * yf).sum(0)


def _finalize(sums, M, C, eps, momentum, running_mean, running_var, stats, update_running):
    mean = sums[:C] / M
    var = (sums[C: 2 * C] / M - mean * mean).clamp_min(0)
    stats[:C] = mean
    stats[C: 2 * C] = torch.rsqrt(var + eps)
    if update_running and running_mean is not None:
        unb = var * (M / (M - 1)) if M > 1 else var
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * unb)


def bn_finalize(sums, M, C, eps, momentum, running_mean=None, running_var=None, stats=None, from_running: bool = False,
                update_running: bool = True):
    if from_running:
        stats[:C] = running_mean[:C]
        stats[C: 2 * C] = torch.rsqrt(running_var[:C] + eps)
        return
    _finalize(sums, M, C, eps, momentum, running_mean, running_var, stats, update_running)
    sums[: 2 * C] = 0


def bn_apply(y, M, C, stats, gamma, beta, z, relu: bool = False, residual=None, sums=None, eps: float = 1e-5, momentum: float = 0.1,
             running_mean=None, running_var=None):
    if sums is not None:
        _finalize(sums, M, C, eps, momentum, running_mean, running_var, stats, running_mean is not None)
    v = (_f(y[:M, :C]) - stats[:C]) * stats[C: 2 * C] * gamma[:C] + beta[:C]
    if residual is not None:
        v = v + _f(residual[:M, :C])
    if relu:
        v = torch.relu(v)
    z[:M, :C] = v.to(z.dtype)


def bn_backward(reduce, y, M, C, stats, gamma, z, relu, dz_a, dz_b=None, dgamma=None, dbeta=None, dy=None, g_out=None):
    g = _f(dz_a[:M, :C])
    if dz_b is not None:
        g = g + _f(dz_b[:M, :C])
    if relu:
        g = g * (_f(z[:M, :C]) > 0)
    xhat = (_f(y[:M, :C]) - stats[:C]) * stats[C: 2 * C]
    if reduce:
        dbeta[:C] += g.sum(0)
        dgamma[:C] += (g * xhat).sum(0)
        return
    d = gamma[:C] * stats[C: 2 * C] * (g - dbeta[:C] / M - xhat * dgamma[:C] / M)
    dy[:M, :C] = d.to(dy.dtype)
    if g_out is not None:
        g_out[:M, :C] = g.to(g_out.dtype)


# ----------------------------------------------------------------------------- pooling / loss
def maxpool_cl(backward, x_or_dx, y_or_dy, idx, B, H, W, C, k, s, p, dy_b=None):
    PH, PW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    if not backward:
        xn = _f(x_or_dx[: B * H * W, :C]).reshape(B, H, W, C).permute(0, 3, 1, 2)
        out, flat = F.max_pool2d(xn, k, s, p, return_indices=True)            # flat index into H*W
        hh, ww = flat // W, flat % W
        ph = torch.arange(PH).view(1, 1, PH, 1)
        pw = torch.arange(PW).view(1, 1, 1, PW)
        tap = (hh - (ph * s - p)) * k + (ww - (pw * s - p))
        y_or_dy[: B * PH * PW, :C] = out.permute(0, 2, 3, 1).reshape(B * PH * PW, C).to(y_or_dy.dtype)
        idx[: B * PH * PW * C] = tap.permute(0, 2, 3, 1).reshape(-1).to(torch.uint8)
        return
    g = _f(y_or_dy[: B * PH * PW, :C])
    if dy_b is not None:
        g = g + _f(dy_b[: B * PH * PW, :C])
    g = g.reshape(B, PH, PW, C)
    tap = idx[: B * PH * PW * C].reshape(B, PH, PW, C).to(torch.int64)
    ph = torch.arange(PH).view(1, PH, 1, 1)
    pw = torch.arange(PW).view(1, 1, PW, 1)
    hh = ph * s - p + tap // k
    ww = pw * s - p + tap % k
    dx = torch.zeros(B, H * W, C)
    lin = (hh * W + ww).reshape(B, PH * PW, C)
    dx.scatter_add_(1, lin, g.reshape(B, PH * PW, C))
    x_or_dx[: B * H * W, :C] = dx.reshape(B * H * W, C).to(x_or_dx.dtype)


def avgpool_cl(backward, x_or_dx, y_or_dy, B, HW, C):
    if not backward:
        y_or_dy[:B, :C] = _f(x_or_dx[: B * HW, :C]).reshape(B, HW, C).mean(1).to(y_or_dy.dtype)
    else:
        x_or_dx[: B * HW, :C] = (_f(y_or_dy[:B, :C]) / HW).repeat_interleave(HW, 0).to(x_or_dx.dtype)


def softmax_ce(logits, labels, C, loss_slot, dz: Optional[torch.Tensor], dzt, db, pad_cols, win_state=None, window_rows: int = 0,
               n_windows: int = 1):
    B = logits.shape[0]
    z = _f(logits[:, :C])
    off = _window_offset(win_state, window_rows, n_windows) if win_state is not None else 0
    lab = labels[off: off + B].to(torch.int64)
    loss_slot[0] += F.cross_entropy(z, lab, reduction="mean")
    if dz is not None or dzt is not None or db is not None:
        g = (torch.softmax(z, 1) - F.one_hot(lab, C)) / B
        if dz is not None:
            dz[:, :C] = g.to(dz.dtype)
        if dzt is not None:
            dzt[:C, :B] = g.t().to(dzt.dtype)
        if db is not None:
            db[:C] += g.sum(0)


def mse_loss(pred, target, D, loss_slot, dz, dzt, db, pad_cols, win_state=None, window_rows: int = 0, n_windows: int = 1):
    B = pred.shape[0]
    off = _window_offset(win_state, window_rows, n_windows) if win_state is not None else 0
    dlt = _f(pred[:, :D]) - _f(target[off: off + B, :D])
    loss_slot[0] += (dlt * dlt).mean()
    g = 2.0 * dlt / (B * D)
    if dz is not None:
        dz[:, :D] = g.to(dz.dtype)
    if dzt is not None:
        dzt[:D, :B] = g.t().to(dzt.dtype)
    if db is not None:
        db[:D] += g.sum(0)


def dropout2d_keep_mask(seed: int, iteration: int, B: int, C: int, p_drop: float) -> torch.Tensor:
    """The kernel's counter-based channel dropout (csrc/conv.cu ``hash3``): [B, C] bool, True = channel kept.
    A pure function of (seed, iteration, sample, channel), so forward and backward agree without storing a mask and a
    test can apply the very same mask to an autograd model."""
    import numpy as np

    M = 0xFFFFFFFF
    a = np.uint64((int(seed) + int(iteration) * 0x632BE5AB) & M)
    b = np.arange(B, dtype=np.uint64).reshape(B, 1)
    c = np.arange(C, dtype=np.uint64).reshape(1, C)
    x = ((a * np.uint64(0x9E3779B1)) & np.uint64(M)) ^ (((b + np.uint64(0x7F4A7C15)) & np.uint64(M)) * np.uint64(0x85EBCA77) & np.uint64(M)) \
        ^ (((c + np.uint64(0x165667B1)) & np.uint64(M)) * np.uint64(0xC2B2AE3D) & np.uint64(M))
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(M)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(M)
    x ^= x >> np.uint64(16)
    thresh = np.uint64(int(float(p_drop) * 4294967296.0)) if p_drop > 0 else np.uint64(0)
    return torch.from_numpy(x >= thresh)


def maxpool_dropout(backward, x_or_dx, ld_x, y_or_dy, idx, B, H, W, C, k, p_drop: float = 0.0, train: bool = False, seed: int = 0,
                    state=None, y_nhwc: bool = False, relu_x=None, db=None):
    """k x k / stride k max-pool (floor) over NHWC, then dropout2d (whole channels of a sample zeroed, survivors scaled by
    1/(1-p)); the pooled map is written NHWC or in torch's flatten order (c, ph, pw)."""
    drop_on = bool(train and p_drop > 0)
    keep = None
    if drop_on:
        it = (int(state[4]) + (int(state[5]) << 32)) if state is not None else 0
        keep = dropout2d_keep_mask(seed, it, B, C, p_drop).view(B, 1, 1, C).to(torch.float32)
    keep_scale = 1.0 / (1.0 - p_drop) if drop_on else 1.0
    PH, PW = H // k, W // k
    if not backward:
        xn = _f(x_or_dx[: B * H * W, :C]).reshape(B, H, W, C)[:, : PH * k, : PW * k]
        win = xn.reshape(B, PH, k, PW, k, C).permute(0, 1, 3, 5, 2, 4).reshape(B, PH, PW, C, k * k)
        best, arg = win.max(-1)
        if drop_on:
            best = best * keep * keep_scale
        if y_nhwc:
            y_or_dy[: B * PH * PW, :C] = best.reshape(B * PH * PW, C).to(y_or_dy.dtype)
        else:
            y_or_dy[:B, : C * PH * PW] = best.permute(0, 3, 1, 2).reshape(B, C * PH * PW).to(y_or_dy.dtype)
        idx[: B * PH * PW * C] = arg.reshape(-1).to(torch.uint8)
        return
    if y_nhwc:
        g = _f(y_or_dy[: B * PH * PW, :C]).reshape(B, PH, PW, C)
    else:
        g = _f(y_or_dy[:B, : C * PH * PW]).reshape(B, C, PH, PW).permute(0, 2, 3, 1)
    arg = idx[: B * PH * PW * C].reshape(B, PH, PW, C).to(torch.int64)
    if drop_on:
        g = g * keep * keep_scale
    onehot = F.one_hot(arg, k * k).to(torch.float32)                                # [B, PH, PW, C, kk]
    if relu_x is not None:
        xr = _f(relu_x[: B * H * W, :C]).reshape(B, H, W, C)[:, : PH * k, : PW * k]
        xw = xr.reshape(B, PH, k, PW, k, C).permute(0, 1, 3, 5, 2, 4).reshape(B, PH, PW, C, k * k)
        g = g * ((xw * onehot).sum(-1) > 0)
    d = (onehot * g.unsqueeze(-1)).reshape(B, PH, PW, C, k, k).permute(0, 1, 4, 2, 5, 3).reshape(B, PH * k, PW * k, C)
    dx = _f(x_or_dx[: B * H * W, :C]).reshape(B, H, W, C).clone()
    dx[:, : PH * k, : PW * k] = d
    x_or_dx[: B * H * W, :C] = dx.reshape(B * H * W, C).to(x_or_dx.dtype)
    if db is not None:
        db[:C] += g.sum((0, 1, 2))
