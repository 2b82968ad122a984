"""Executable model of the implicit-GEMM convolution (csrc/gemm_sm100.cu conv2d_fwd_implicit): the im2col-mode TMA load as its
tensor map is documented (cuda.h cuTensorMapEncodeIm2col; CUTLASS cute/atom/copy_traits_sm90_im2col.hpp): a load visits
`pixels` consecutive positions of the bounding box  [lower, extent + upper)  in (w, h, n) order starting at the receptive-field
corner (w, h) of image n, steps by the traversal stride, adds the filter-tap offset to every position and reads 0 outside the image.
The test runs the kernel's host/device index arithmetic (tile -> first pixel -> corner coordinates, k-block -> tap / channel block,
weight order (r, s, c), corners  lower = -pad, upper = pad - (filter - 1) * dilation) against F.conv2d.  It pins down the
CONVENTIONS the kernel was written to; the kernel itself has not run on a GPU yet."""
import pytest
import torch
import torch.nn.functional as F

BM, BK = 128, 64


def tma_im2col_load(x, c0, w, h, n, off_w, off_h, lower, upper, stride, pixels=BM, channels=BK):
    N, H, W, C = x.shape
    out = torch.zeros(pixels, channels)
    w_lo, h_lo = lower
    w_hi, h_hi = W + upper[0], H + upper[1]              # exclusive bounds of the corner positions
    for p in range(pixels):
        if n < N:
            ww, hh = w + off_w, h + off_h
            if 0 <= ww < W and 0 <= hh < H:
                out[p] = x[n, hh, ww, c0:c0 + channels]
        w += stride
        if w >= w_hi:
            w = w_lo
            h += stride
            if h >= h_hi:
                h = h_lo
                n += 1
    return out


def conv_implicit_model(x, wk, R, S, stride, pad, dil):
    N, H, W, C = x.shape
    Cout = wk.shape[0]
    P = (H + 2 * pad - dil * (R - 1) - 1) // stride + 1
    Q = (W + 2 * pad - dil * (S - 1) - 1) // stride + 1
    lower, upper = (-pad, -pad), (pad - (S - 1) * dil, pad - (R - 1) * dil)
    M, cb_n = N * P * Q, C // BK
    out = torch.zeros(((M + BM - 1) // BM) * BM, Cout)
    for m0 in range(0, M, BM):                                             # one CTA tile
        img, rem = divmod(m0, P * Q)
        op, oq = divmod(rem, Q)
        acc = torch.zeros(BM, Cout)
        for kb in range(R * S * cb_n):                                     # one k-block = one filter tap x 64 channels
            tap, cb = divmod(kb, cb_n)
            fr, fs = divmod(tap, S)
            a = tma_im2col_load(x, cb * BK, oq * stride - pad, op * stride - pad, img, fs * dil, fr * dil, lower, upper, stride)
            acc += a @ wk[:, kb * BK:(kb + 1) * BK].t()
        out[m0:m0 + BM] = acc
    return out[:M].view(N, P, Q, Cout)


@pytest.mark.parametrize("N,H,W,C,Cout,k,stride,pad,dil", [(2, 8, 8, 64, 8, 3, 1, 1, 1), (3, 10, 7, 64, 5, 3, 1, 1, 1), (2, 16, 16, 64, 4, 3, 2, 1, 1),
                                                           (2, 9, 9, 128, 4, 3, 1, 0, 1), (1, 12, 12, 64, 4, 1, 1, 0, 1), (2, 11, 11, 64, 4, 3, 1, 2, 2)])
def test_index_arithmetic_of_the_implicit_gemm_convolution(N, H, W, C, Cout, k, stride, pad, dil):
    torch.manual_seed(0)
    x = torch.randn(N, C, H, W)
    w = torch.randn(Cout, C, k, k) * 0.1
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, k * k * C)                    # reduction index ordered (r, s, c)
    got = conv_implicit_model(xn, wk, k, k, stride, pad, dil)
    ref = F.conv2d(x, w, None, stride=stride, padding=pad, dilation=dil).permute(0, 2, 3, 1)
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
