"""csrc/norm.cu kernels against plain PyTorch fp32 references of the same ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from sparktorch_b200.ops import native

    return native.load(required=True)


def _dev():
    return torch.device("cuda:0")


def _nhwc(x):            # [B, C, H, W] fp32 -> [B*H*W, C] bf16
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous().to(torch.bfloat16)


def _nchw(m, b, h, w):   # [B*H*W, C] -> [B, C, H, W] fp32
    return m.float().reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("k,s,p,dil,C_", [(3, 1, 1, 1, 16), (3, 2, 1, 1, 64), (1, 2, 0, 1, 32), (3, 1, 1, 2, 16), (1, 1, 0, 2, 8)])
def test_im2col_cl_matches_unfold(C, k, s, p, dil, C_):
    dev = _dev()
    B, H, W = 3, 9, 7
    x = torch.randn(B, C_, H, W, device=dev).to(torch.bfloat16).float()
    xd = x
    if dil > 1:          # zero insertion = what a strided convolution's dgrad correlates against
        xd = torch.zeros(B, C_, (H - 1) * dil + 1, (W - 1) * dil + 1, device=dev)
        xd[:, :, ::dil, ::dil] = x
    Hd, Wd = xd.shape[2:]
    OH, OW = (Hd + 2 * p - k) // s + 1, (Wd + 2 * p - k) // s + 1
    ref = F.unfold(xd, k, padding=p, stride=s)                       # [B, C*k*k, OH*OW], (c, i, j) rows
    ref = ref.reshape(B, C_, k * k, OH * OW).permute(0, 3, 2, 1).reshape(B * OH * OW, k * k * C_)   # -> (i, j, c)
    col = torch.full((B * OH * OW, k * k * C_), 7.0, dtype=torch.bfloat16, device=dev)
    C.im2col_cl(_nhwc(x), B, H, W, C_, k, k, s, p, dil, OH, OW, col)
    torch.testing.assert_close(col.float(), ref, rtol=0, atol=0)


def test_flip_conv_weight_cl(C):
    dev = _dev()
    Cout, Cin, k = 40, 24, 3
    w = torch.randn(Cout, Cin, k, k, device=dev).to(torch.bfloat16)
    w_cl = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
    out = torch.zeros(Cin, k * k * Cout, dtype=torch.bfloat16, device=dev)
    C.flip_conv_weight_cl(w_cl, Cout, Cin, k, k, out)
    ref = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, k * k * Cout)     # [ci, (i, j), co] of the flipped taps
    torch.testing.assert_close(out.float(), ref.float(), rtol=0, atol=0)


@pytest.mark.parametrize("C_,M,relu,res", [(64, 1000, True, False), (256, 333, True, True), (2048, 40, False, False), (16, 4096, True, True)])
def test_batchnorm_forward_backward(C, C_, M, relu, res):
    dev = _dev()
    torch.manual_seed(0)
    y = (torch.randn(M, C_, device=dev) * 2 + 0.5).to(torch.bfloat16)
    r = torch.randn(M, C_, device=dev).to(torch.bfloat16) if res else None
    gamma = torch.rand(C_, device=dev) + 0.5
    beta = torch.randn(C_, device=dev)
    rm, rv = torch.zeros(C_, device=dev), torch.ones(C_, device=dev)
    sums = torch.zeros(2 * C_, device=dev)
    stats = torch.zeros(2 * C_, device=dev)
    z = torch.zeros(M, C_, dtype=torch.bfloat16, device=dev)
    C.bn_stats(y, M, C_, sums)
    C.bn_finalize(sums, M, C_, 1e-5, 0.1, rm, rv, stats)
    C.bn_apply(y, M, C_, stats, gamma, beta, z, relu=relu, residual=r)
    # reference (fp32 math on the same bf16 inputs)
    yf = y.float().requires_grad_(True)
    gf, bf_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C_, device=dev), torch.ones(C_, device=dev)
    out = F.batch_norm(yf, rm2, rv2, gf, bf_, training=True, momentum=0.1, eps=1e-5)
    rf = r.float().requires_grad_(True) if res else None
    if res:
        out = out + rf
    if relu:
        out = torch.relu(out)
    torch.testing.assert_close(z.float(), out.detach(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(rm, rm2, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(rv, rv2, rtol=1e-3, atol=1e-4)
    assert float(sums.abs().max()) == 0.0                      # accumulators cleared for the next step
    # backward with two incoming gradients (residual join)
    da = torch.randn(M, C_, device=dev).to(torch.bfloat16)
    db = torch.randn(M, C_, device=dev).to(torch.bfloat16)
    # the kernel takes its ReLU mask from the stored bf16 output; use exactly that mask in the reference
    mask = (z.float() > 0).float() if relu else torch.ones(M, C_, device=dev)
    pre = F.batch_norm(yf, None, None, gf, bf_, training=True, eps=1e-5)
    g_in = (da.float() + db.float()) * mask
    (pre * g_in).sum().backward()
    dgamma, dbeta = torch.zeros(C_, device=dev), torch.zeros(C_, device=dev)
    dy = torch.zeros(M, C_, dtype=torch.bfloat16, device=dev)
    gout = torch.zeros(M, C_, dtype=torch.bfloat16, device=dev)
    C.bn_backward(True, y, M, C_, stats, gamma, z, relu, da, dz_b=db, dgamma=dgamma, dbeta=dbeta)
    C.bn_backward(False, y, M, C_, stats, gamma, z, relu, da, dz_b=db, dgamma=dgamma, dbeta=dbeta, dy=dy, g_out=gout)
    torch.testing.assert_close(dgamma, gf.grad, rtol=2e-3, atol=2e-2)
    torch.testing.assert_close(dbeta, bf_.grad, rtol=2e-3, atol=2e-2)
    torch.testing.assert_close(gout.float(), g_in, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dy.float(), yf.grad, rtol=2e-2, atol=2e-2)


def test_batchnorm_eval_uses_running_stats(C):
    dev = _dev()
    M, C_ = 64, 32
    y = torch.randn(M, C_, device=dev).to(torch.bfloat16)
    rm, rv = torch.randn(C_, device=dev), torch.rand(C_, device=dev) + 0.5
    gamma, beta = torch.rand(C_, device=dev) + 0.5, torch.randn(C_, device=dev)
    stats, sums = torch.zeros(2 * C_, device=dev), torch.zeros(2 * C_, device=dev)
    z = torch.zeros(M, C_, dtype=torch.bfloat16, device=dev)
    rm0, rv0 = rm.clone(), rv.clone()
    C.bn_finalize(sums, M, C_, 1e-5, 0.1, rm, rv, stats, from_running=True, update_running=False)
    C.bn_apply(y, M, C_, stats, gamma, beta, z)
    ref = F.batch_norm(y.float(), rm0, rv0, gamma, beta, training=False, eps=1e-5)
    torch.testing.assert_close(z.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(rm, rm0, rtol=0, atol=0)


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (2, 2, 0), (3, 1, 1)])
def test_maxpool_cl_forward_backward(C, k, s, p):
    dev = _dev()
    B, C_, H, W = 2, 16, 11, 9
    x = torch.randn(B, C_, H, W, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    ref = F.max_pool2d(x, k, s, p)
    PH, PW = ref.shape[2:]
    y = torch.zeros(B * PH * PW, C_, dtype=torch.bfloat16, device=dev)
    idx = torch.zeros(B * PH * PW * C_, dtype=torch.uint8, device=dev)
    C.maxpool_cl(False, _nhwc(x.detach()), y, idx, B, H, W, C_, k, s, p)
    torch.testing.assert_close(_nchw(y, B, PH, PW), ref.detach(), rtol=0, atol=0)
    ga = torch.randn(B, C_, PH, PW, device=dev).to(torch.bfloat16).float()
    gb = torch.randn(B, C_, PH, PW, device=dev).to(torch.bfloat16).float()
    ref.backward(ga + gb)
    dx = torch.full((B * H * W, C_), 3.0, dtype=torch.bfloat16, device=dev)
    C.maxpool_cl(True, dx, _nhwc(ga), idx, B, H, W, C_, k, s, p, dy_b=_nhwc(gb))
    torch.testing.assert_close(_nchw(dx, B, H, W), x.grad, rtol=2e-2, atol=2e-2)


def test_avgpool_cl(C):
    dev = _dev()
    B, HW, C_ = 5, 49, 64
    x = torch.randn(B * HW, C_, device=dev).to(torch.bfloat16)
    y = torch.zeros(B, C_, dtype=torch.bfloat16, device=dev)
    C.avgpool_cl(False, x, y, B, HW, C_)
    torch.testing.assert_close(y.float(), x.float().reshape(B, HW, C_).mean(1), rtol=1e-2, atol=1e-2)
    dy = torch.randn(B, C_, device=dev).to(torch.bfloat16)
    dx = torch.zeros(B * HW, C_, dtype=torch.bfloat16, device=dev)
    C.avgpool_cl(True, dx, dy, B, HW, C_)
    torch.testing.assert_close(dx.float(), (dy.float() / HW).repeat_interleave(HW, 0), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("k,s,p,Cin,Cout", [(3, 1, 1, 16, 32), (3, 2, 1, 32, 32), (1, 2, 0, 32, 64), (1, 1, 0, 64, 16)])
def test_conv_forward_dgrad_wgrad_via_gemm(C, k, s, p, Cin, Cout):
    """The three GEMMs of a channels-last convolution exactly as runtime/fused_resnet.py issues them."""
    dev = _dev()
    B, H, W = 4, 12, 10
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, Cin, H, W, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, device=dev) * 0.2).to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv2d(x, w, stride=s, padding=p)
    dy_ref = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(dy_ref)
    X = _nhwc(x.detach())
    Wb = w.detach().permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous().to(torch.bfloat16)
    M, K = B * OH * OW, k * k * Cin
    if k == 1 and s == 1:
        A = X
    else:
        A = torch.zeros(M, K, dtype=torch.bfloat16, device=dev)
        C.im2col_cl(X, B, H, W, Cin, k, k, s, p, 1, OH, OW, A)
    Y = torch.zeros(M, Cout, dtype=torch.bfloat16, device=dev)
    C.gemm_bf16(A, False, Wb, False, M, Cout, K, out_bf16=Y)
    torch.testing.assert_close(_nchw(Y, B, OH, OW), ref.detach(), rtol=2e-2, atol=5e-2)
    dY = _nhwc(dy_ref)
    gW = torch.zeros(Cout, K, device=dev)
    C.gemm_bf16(dY, True, A, True, Cout, K, M, out_f32=gW, atomic_f32=True, split_k=3)
    gw_ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, K)
    torch.testing.assert_close(gW, gw_ref, rtol=2e-2, atol=0.15)
    gX = torch.zeros(B * H * W, Cin, dtype=torch.bfloat16, device=dev)
    if k == 1 and s == 1:
        C.gemm_bf16(dY, False, Wb, True, M, Cin, Cout, out_bf16=gX)
    elif k == 1:
        gXc = torch.zeros(M, Cin, dtype=torch.bfloat16, device=dev)
        C.gemm_bf16(dY, False, Wb, True, M, Cin, Cout, out_bf16=gXc)
        C.im2col_cl(gXc, B, OH, OW, Cin, 1, 1, 1, 0, s, H, W, gX)
    else:
        Wf = torch.zeros(Cin, k * k * Cout, dtype=torch.bfloat16, device=dev)
        C.flip_conv_weight_cl(Wb, Cout, Cin, k, k, Wf)
        colT = torch.zeros(B * H * W, k * k * Cout, dtype=torch.bfloat16, device=dev)
        C.im2col_cl(dY, B, OH, OW, Cout, k, k, 1, k - 1 - p, s, H, W, colT)
        C.gemm_bf16(colT, False, Wf, False, B * H * W, Cin, k * k * Cout, out_bf16=gX)
    torch.testing.assert_close(_nchw(gX, B, H, W), x.grad, rtol=2e-2, atol=0.1)


@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (20000, 256, 64), (333, 128, 576)])
def test_gemm_epilogue_batch_stats_feed_bn_apply(C, M, N, K):
    """conv GEMM with colsum/colsumsq epilogue + bn_apply(inline finalize) == bn_stats -> bn_finalize -> bn_apply."""
    dev = _dev()
    torch.manual_seed(1)
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.2).to(torch.bfloat16)
    gamma, beta = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev)
    Y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    sums = torch.zeros(2 * N, device=dev)
    C.gemm_bf16(A, False, W, False, M, N, K, out_bf16=Y, colsum=sums[:N], colsumsq=sums[N:])
    yf = A.float() @ W.float().t()
    torch.testing.assert_close(sums[:N], yf.sum(0), rtol=2e-3, atol=0.5)
    torch.testing.assert_close(sums[N:], (yf * yf).sum(0), rtol=2e-3, atol=0.5)
    stats = torch.zeros(2 * N, device=dev)
    rm, rv = torch.zeros(N, device=dev), torch.ones(N, device=dev)
    Z = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    C.bn_apply(Y, M, N, stats, gamma, beta, Z, relu=True, sums=sums, eps=1e-5, momentum=0.1, running_mean=rm, running_var=rv)
    rm2, rv2 = torch.zeros(N, device=dev), torch.ones(N, device=dev)
    ref = torch.relu(F.batch_norm(yf, rm2, rv2, gamma, beta, training=True, momentum=0.1, eps=1e-5))
    torch.testing.assert_close(Z.float(), ref, rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(rm, rm2, rtol=1e-2, atol=1e-3)
    torch.testing.assert_close(rv, rv2, rtol=1e-2, atol=1e-3)
    torch.testing.assert_close(stats[:N], yf.mean(0), rtol=1e-2, atol=1e-3)
