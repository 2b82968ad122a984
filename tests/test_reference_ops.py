"""ops/reference.py is the oracle for the kernels and the back-end of the plan emulation: check IT against
torch.nn.functional (convolution as im2col + GEMM in all three directions, batch norm, pooling, losses)."""
import pytest
import torch
import torch.nn.functional as F

from sparktorch_b200.ops import reference as R


def _nhwc(x):
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous()


def _nchw(m, b, h, w):
    return m.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("k,s,p,Cin,Cout", [(3, 1, 1, 8, 16), (3, 2, 1, 16, 16), (1, 2, 0, 8, 24), (1, 1, 0, 16, 8), (5, 1, 2, 8, 8)])
def test_channels_last_convolution_three_gemms(k, s, p, Cin, Cout):
    torch.manual_seed(0)
    B, H, W = 3, 9, 8
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, Cin, H, W, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, requires_grad=True)
    ref = F.conv2d(x, w, stride=s, padding=p)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    X, Wb = _nhwc(x.detach()), w.detach().permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
    M, K = B * OH * OW, k * k * Cin
    A = torch.zeros(M, K)
    if k == 1 and s == 1:
        A = X
    else:
        R.im2col_cl(X, B, H, W, Cin, k, k, s, p, 1, OH, OW, A)
    Y = torch.zeros(M, Cout)
    R.gemm_bf16(A, False, Wb, False, M, Cout, K, out_bf16=Y)
    torch.testing.assert_close(_nchw(Y, B, OH, OW), ref.detach(), rtol=1e-4, atol=1e-4)
    dY = _nhwc(dy)
    gW = torch.zeros(Cout, K)
    R.gemm_bf16(dY, True, A, True, Cout, K, M, out_f32=gW, atomic_f32=True, split_k=4)
    torch.testing.assert_close(gW, w.grad.permute(0, 2, 3, 1).reshape(Cout, K), rtol=1e-4, atol=1e-4)
    gX = torch.zeros(B * H * W, Cin)
    if k == 1 and s == 1:
        R.gemm_bf16(dY, False, Wb, True, M, Cin, Cout, out_bf16=gX)
    elif k == 1:
        gXc = torch.zeros(M, Cin)
        R.gemm_bf16(dY, False, Wb, True, M, Cin, Cout, out_bf16=gXc)
        R.im2col_cl(gXc, B, OH, OW, Cin, 1, 1, 1, 0, s, H, W, gX)
    else:
        Wf = torch.zeros(Cin, k * k * Cout)
        R.flip_conv_weight_cl(Wb, Cout, Cin, k, k, Wf)
        colT = torch.zeros(B * H * W, k * k * Cout)
        R.im2col_cl(dY, B, OH, OW, Cout, k, k, 1, k - 1 - p, s, H, W, colT)
        R.gemm_bf16(colT, False, Wf, False, B * H * W, Cin, k * k * Cout, out_bf16=gX)
    torch.testing.assert_close(_nchw(gX, B, H, W), x.grad, rtol=1e-4, atol=1e-4)


def test_generic_stride_im2col_and_flipped_taps_match_torch_layouts():
    torch.manual_seed(1)
    B, C, H, W, k = 2, 3, 7, 6, 3
    x = torch.randn(B, C * H * W)                            # Spark rows: (c, h, w) order
    col = torch.zeros(B * (H - 2) * (W - 2), C * k * k)
    R.im2col_nhwc(x, B, H, W, C, x.stride(0), W, 1, H * W, k, k, 1, 0, col)
    ref = F.unfold(x.view(B, C, H, W), k).permute(0, 2, 1).reshape(-1, C * k * k)
    torch.testing.assert_close(col, ref)
    w = torch.randn(5, C, k, k)
    out = torch.zeros(C, 5 * k * k)
    R.flip_conv_weight(w.reshape(5, -1), 5, C, k, k, out)
    torch.testing.assert_close(out, w.flip(2, 3).permute(1, 0, 2, 3).reshape(C, -1))


@pytest.mark.parametrize("relu,res", [(True, True), (False, False)])
def test_batchnorm_forward_backward(relu, res):
    torch.manual_seed(2)
    M, C = 50, 16
    y = torch.randn(M, C) * 2 + 0.3
    r = torch.randn(M, C) if res else None
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C)
    rm, rv = torch.zeros(C), torch.ones(C)
    sums, stats, z = torch.zeros(2 * C), torch.zeros(2 * C), torch.zeros(M, C)
    R.bn_stats(y, M, C, sums)
    R.bn_finalize(sums, M, C, 1e-5, 0.1, rm, rv, stats)
    R.bn_apply(y, M, C, stats, gamma, beta, z, relu=relu, residual=r)
    yf, gf, bf = y.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C), torch.ones(C)
    out = F.batch_norm(yf, rm2, rv2, gf, bf, training=True, momentum=0.1, eps=1e-5)
    if res:
        out = out + r
    if relu:
        out = torch.relu(out)
    torch.testing.assert_close(z, out.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rm, rm2, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rv, rv2, rtol=1e-4, atol=1e-6)
    da, db = torch.randn(M, C), torch.randn(M, C)
    out.backward(da + db)
    dgamma, dbeta, dy, gout = torch.zeros(C), torch.zeros(C), torch.zeros(M, C), torch.zeros(M, C)
    R.bn_backward(True, y, M, C, stats, gamma, z, relu, da, dz_b=db, dgamma=dgamma, dbeta=dbeta)
    R.bn_backward(False, y, M, C, stats, gamma, z, relu, da, dz_b=db, dgamma=dgamma, dbeta=dbeta, dy=dy, g_out=gout)
    torch.testing.assert_close(dgamma, gf.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbeta, bf.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dy, yf.grad, rtol=1e-3, atol=1e-5)
    # in-epilogue statistics variant: sums handed to bn_apply
    sums2, stats2, z2 = torch.zeros(2 * C), torch.zeros(2 * C), torch.zeros(M, C)
    R.bn_stats(y, M, C, sums2)
    R.bn_apply(y, M, C, stats2, gamma, beta, z2, relu=relu, residual=r, sums=sums2, eps=1e-5, momentum=0.1)
    torch.testing.assert_close(z2, z)
    torch.testing.assert_close(stats2, stats)


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (2, 2, 0)])
def test_pooling(k, s, p):
    torch.manual_seed(3)
    B, C, H, W = 2, 8, 9, 7
    x = torch.randn(B, C, H, W, requires_grad=True)
    ref = F.max_pool2d(x, k, s, p)
    PH, PW = ref.shape[2:]
    y, idx = torch.zeros(B * PH * PW, C), torch.zeros(B * PH * PW * C, dtype=torch.uint8)
    R.maxpool_cl(False, _nhwc(x.detach()), y, idx, B, H, W, C, k, s, p)
    torch.testing.assert_close(_nchw(y, B, PH, PW), ref.detach())
    ga, gb = torch.randn_like(ref), torch.randn_like(ref)
    ref.backward(ga + gb)
    dx = torch.zeros(B * H * W, C)
    R.maxpool_cl(True, dx, _nhwc(ga), idx, B, H, W, C, k, s, p, dy_b=_nhwc(gb))
    torch.testing.assert_close(_nchw(dx, B, H, W), x.grad, rtol=1e-5, atol=1e-6)
    # k x k / stride k pool of the CNN plan, (c, ph, pw) flatten order, ReLU mask + bias gradient in the backward
    xr = torch.relu(torch.randn(B, C, 6, 6)).requires_grad_(True)
    pooled = F.max_pool2d(xr, 2)
    flat, idx2 = torch.zeros(B, C * 9), torch.zeros(B * 9 * C, dtype=torch.uint8)
    R.maxpool_dropout(False, _nhwc(xr.detach()), C, flat, idx2, B, 6, 6, C, 2)
    torch.testing.assert_close(flat, torch.flatten(pooled, 1).detach())
    g = torch.randn(B, C * 9)
    torch.flatten(pooled, 1).backward(g)
    dx2, dbias = torch.zeros(B * 36, C), torch.zeros(C)
    R.maxpool_dropout(True, dx2, C, g, idx2, B, 6, 6, C, 2, relu_x=_nhwc(xr.detach()), db=dbias)
    want = xr.grad * (xr.detach() > 0)
    torch.testing.assert_close(_nchw(dx2, B, 6, 6), want, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dbias, want.sum((0, 2, 3)), rtol=1e-5, atol=1e-5)
    # global average pool
    a = torch.randn(B * 12, C)
    out = torch.zeros(B, C)
    R.avgpool_cl(False, a, out, B, 12, C)
    torch.testing.assert_close(out, a.reshape(B, 12, C).mean(1))


def test_losses_fused_and_standalone_agree():
    torch.manual_seed(4)
    Bt, K, N = 12, 7, 5
    A, Wt, bias = torch.randn(Bt, K), torch.randn(N, K), torch.randn(N)
    y = torch.randint(0, N, (Bt,), dtype=torch.int32)
    z = (A @ Wt.t() + bias).requires_grad_(True)
    ce = F.cross_entropy(z, y.long())
    ce.backward()
    loss, dz, db, raw = torch.zeros(1), torch.zeros(Bt, N), torch.zeros(N), torch.zeros(Bt, N)
    R.gemm_bf16(A, False, Wt, False, Bt, N, K, bias=bias, loss_kind=1, labels=y, loss_slot=loss, loss_scale=1.0 / Bt, out_bf16=dz,
                colsum=db, out_f32=raw)
    torch.testing.assert_close(loss[0], ce.detach())
    torch.testing.assert_close(dz, z.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(db, z.grad.sum(0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(raw, z.detach())
    l2, dz2, db2 = torch.zeros(1), torch.zeros(Bt, N), torch.zeros(N)
    R.softmax_ce(z.detach(), y, N, l2, dz2, None, db2, 8)
    torch.testing.assert_close(l2, loss)
    torch.testing.assert_close(dz2, dz)
    t = torch.randn(Bt, N)
    l3, dz3 = torch.zeros(1), torch.zeros(Bt, N)
    R.gemm_bf16(A, False, Wt, False, Bt, N, K, bias=bias, loss_kind=2, target=t, loss_slot=l3, loss_scale=1.0 / (Bt * N), out_bf16=dz3)
    l4, dz4 = torch.zeros(1), torch.zeros(Bt, N)
    R.mse_loss(z.detach(), t, N, l4, dz4, None, None, 8)
    torch.testing.assert_close(l3, l4)
    torch.testing.assert_close(dz3, dz4)
    torch.testing.assert_close(l4[0], F.mse_loss(z.detach(), t))
