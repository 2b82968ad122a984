// Batch-norm, channels-last im2col and pooling kernels for residual conv nets
// (SURVEY.md §2.7 G6 at ResNet-50 shapes, G13 BatchNorm).
//
// All activations are channels-last matrices [M = B*H*W, C] in bf16 with C % 8 == 0, so every
// thread moves 16-byte vectors of 8 channels and the same buffers are the A / B operands of the
// tcgen05 GEMM (gemm_sm100.cu) without any repacking:
//
//   conv (1x1)       Y = X * W^T                                         GEMM only
//   conv (kxk, s, p) Y = im2col_cl(X) * W^T, W stored [Cout, kh, kw, Cin] (torch channels_last weight)
//   BN forward       bn_stats (per-channel sum / sum of squares, fp32 red.add) -> bn_finalize (mean, rstd,
//                    running statistics) -> bn_apply (normalise, scale/shift, + residual, ReLU) in one pass
//   BN backward      bn_bwd_reduce (dgamma, dbeta straight into the gradient arena of K-AR, ReLU mask and the
//                    sum of the two incoming gradients of a residual join fused) -> bn_bwd_apply (dY, and the
//                    masked gradient for the skip connection)
//   conv dgrad       im2col_cl of the zero-dilated, padded dY  x  flipped taps (no scatter, no atomics)
//
// The reference runs cuDNN convolutions and ATen batch-norm / pooling kernels for this (it has no model of
// its own here: the config is named in BASELINE.json).
#include "common.cuh"
#include "kernels.h"

namespace st {

constexpr int NT = 256;

struct bf8 { uint4 raw; };
ST_DEVINL void unpack8(const uint4& r, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
ST_DEVINL uint4 pack8(const float* f) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}

// ------------------------------------------------------------------------------------------------ im2col
// col[m, (i*kw + j)*C + c] = x[b, (oh*stride - pad + i)/dil, (ow*stride - pad + j)/dil, c]
// (0 when outside the image or, with input dilation `dil`, between the samples of the dilated map).
__global__ void __launch_bounds__(NT) im2col_cl_kernel(const Im2colClArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  // the kernel is pure address arithmetic + 16-byte copies: keep the arithmetic in 32 bits (the host checks the range)
  const uint32_t c8n = uint32_t(a.C) >> 3;
  const uint32_t kk = uint32_t(a.kh * a.kw);
  const uint32_t per_row = kk * c8n;                                  // 16-byte vectors per patch-matrix row
  const uint32_t M = uint32_t(a.B) * uint32_t(a.OH) * uint32_t(a.OW);
  const uint32_t total = M * per_row;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t e = blockIdx.x * NT + threadIdx.x; e < total; e += gridDim.x * NT) {
    const uint32_t m = e / per_row;
    const uint32_t r = e - m * per_row;
    const uint32_t tap = r / c8n;
    const uint32_t c8 = r - tap * c8n;
    const int i = int(tap / uint32_t(a.kw)), j = int(tap) - i * a.kw;
    const uint32_t t = m / uint32_t(a.OW);
    const int ow = int(m - t * uint32_t(a.OW));
    const uint32_t b = t / uint32_t(a.OH);
    const int oh = int(t - b * uint32_t(a.OH));
    int h = oh * a.stride - a.pad + i, w = ow * a.stride - a.pad + j;
    bool ok = h >= 0 && w >= 0;
    if (a.dil > 1) {
      ok = ok && (h % a.dil == 0) && (w % a.dil == 0);
      h /= a.dil; w /= a.dil;
    }
    ok = ok && h < a.H && w < a.W;
    uint4 v = zero;
    if (ok) v = *reinterpret_cast<const uint4*>(a.x + (int64_t(b) * a.H * a.W + int64_t(h) * a.W + w) * a.ld_x + (c8 << 3));
    *reinterpret_cast<uint4*>(a.col + int64_t(m) * a.ld_col + (r << 3)) = v;
  }
}

static unsigned grid_for_total(int64_t total, int per_sm = 8) {
  int64_t g = (total + NT - 1) / NT;
  if (g < 1) g = 1;
  if (g > 148 * per_sm) g = 148 * per_sm;
  return unsigned(g);
}

int im2col_cl(const Im2colClArgs& a, cudaStream_t stream) {
  if ((a.C & 7) || (a.ld_x & 7) || (a.ld_col & 7)) return -1;
  const int64_t total = int64_t(a.B) * a.OH * a.OW * a.kh * a.kw * (a.C >> 3);
  if (total >= (int64_t(1) << 31)) return -2;
  return int(launch_kernel(im2col_cl_kernel, dim3(grid_for_total(total, 16)), dim3(NT), 0, stream, a));
}

// w [Cout][kh*kw][Cin] (row pitch ld_w) -> out [Cin][kh*kw][Cout] (row pitch ld_o) with flipped taps.
// 32x32 shared-memory tile transpose per tap so both the reads (along ci) and the writes (along co) coalesce.
__global__ void __launch_bounds__(NT) flip_conv_weight_cl_kernel(const __nv_bfloat16* __restrict__ w, int64_t ld_w, int Cout, int Cin,
                                                                 int kk, __nv_bfloat16* __restrict__ out, int64_t ld_o) {
  __shared__ __nv_bfloat16 tile[32][33];
  pdl_wait();
  pdl_launch_dependents();
  const int tap = blockIdx.z;
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + tx;
    tile[r][tx] = (co < Cout && ci < Cin) ? w[int64_t(co) * ld_w + int64_t(tap) * Cin + ci] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + tx;
    if (ci < Cin && co < Cout) out[int64_t(ci) * ld_o + int64_t(kk - 1 - tap) * Cout + co] = tile[tx][r];
  }
}

int flip_conv_weight_cl(const __nv_bfloat16* w, int64_t ld_w, int Cout, int Cin, int kh, int kw, __nv_bfloat16* out, int64_t ld_o,
                        cudaStream_t stream) {
  dim3 grid(unsigned((Cin + 31) / 32), unsigned((Cout + 31) / 32), unsigned(kh * kw));
  return int(launch_kernel(flip_conv_weight_cl_kernel, grid, dim3(NT), 0, stream, w, ld_w, Cout, Cin, kh * kw, out, ld_o));
}

// ------------------------------------------------------------------------------------------------ batch norm
// Thread layout shared by the two column reductions: thread -> (channel group cg of 8 channels, row lane rl);
// a block walks rows [row0, row0 + rows_per_block) with stride `lanes`, partials meet in shared memory and one
// red.add per channel per block goes to global memory.
struct ColLayout { int c8n, lanes, cg, rl; };
ST_DEVINL ColLayout col_layout(int C) {
  ColLayout l;
  l.c8n = C >> 3;
  l.lanes = NT / l.c8n < 1 ? 1 : NT / l.c8n;
  l.cg = threadIdx.x % l.c8n;
  l.rl = threadIdx.x / l.c8n;
  return l;
}

// acc0/acc1: 8 partial sums per thread -> out0[c], out1[c]
ST_DEVINL void block_col_reduce(const ColLayout& l, float* acc0, float* acc1, float* out0, float* out1, float* smem) {
  // smem: [2][lanes][c8n*8]
  const int C = l.c8n << 3;
  if (l.rl < l.lanes) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      smem[(0 * l.lanes + l.rl) * C + (l.cg << 3) + q] = acc0[q];
      smem[(1 * l.lanes + l.rl) * C + (l.cg << 3) + q] = acc1[q];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += NT) {
    const int which = c / C, ch = c - which * C;
    float s = 0.f;
    for (int r = 0; r < l.lanes; ++r) s += smem[(which * l.lanes + r) * C + ch];
    red_add_f32((which ? out1 : out0) + ch, s);
  }
}

// sums[c] += sum_m y[m, c];  sums[C + c] += sum_m y[m, c]^2
__global__ void __launch_bounds__(NT) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, int64_t ld, int64_t M, int C,
                                                      int64_t rows_per_block, float* __restrict__ sums) {
  extern __shared__ float smem[];
  pdl_wait();
  pdl_launch_dependents();
  const ColLayout l = col_layout(C);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  if (l.rl < l.lanes)
    for (int cgi = l.cg; cgi < l.c8n; cgi += NT)          // c8n <= NT: one trip
      for (int64_t m = r0 + l.rl; m < r1; m += l.lanes) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(y + m * ld + (cgi << 3)), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
      }
  block_col_reduce(l, s, q, sums, sums + C, smem);
}

static int64_t rows_per_block_for(int64_t M, int C) {
  const int lanes = NT / (C >> 3) < 1 ? 1 : NT / (C >> 3);
  int64_t rpb = (M + 148 * 4 - 1) / (148 * 4);
  if (rpb < lanes * 4) rpb = lanes * 4;
  return rpb;
}
static size_t col_smem(int C) {
  const int lanes = NT / (C >> 3) < 1 ? 1 : NT / (C >> 3);
  return size_t(2) * lanes * C * sizeof(float);
}

int bn_stats(const __nv_bfloat16* y, int64_t ld, int64_t M, int C, float* sums, cudaStream_t stream) {
  if ((C & 7) || C > 8 * NT || (ld & 7)) return -1;
  const int64_t rpb = rows_per_block_for(M, C);
  const unsigned grid = unsigned((M + rpb - 1) / rpb);
  return int(launch_kernel(bn_stats_kernel, dim3(grid), dim3(NT), col_smem(C), stream, y, ld, M, C, rpb, sums));
}

// stats[c] = mean, stats[C + c] = 1/sqrt(var + eps); running statistics updated like nn.BatchNorm2d in train
// mode (unbiased variance, momentum); the accumulators are cleared for the next step.
// from_running != 0: evaluation mode - stats come from the running buffers, nothing is updated.
__global__ void __launch_bounds__(NT) bn_finalize_kernel(float* __restrict__ sums, int64_t M, int C, float eps, float momentum,
                                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                                         float* __restrict__ stats, int from_running, int update_running) {
  pdl_wait();
  pdl_launch_dependents();
  for (int c = blockIdx.x * NT + threadIdx.x; c < C; c += gridDim.x * NT) {
    float mean, var;
    if (from_running) {
      mean = running_mean[c];
      var = running_var[c];
    } else {
      const float inv = 1.f / float(M);
      mean = sums[c] * inv;
      var = fmaxf(sums[C + c] * inv - mean * mean, 0.f);
      sums[c] = 0.f;
      sums[C + c] = 0.f;
      if (update_running && running_mean != nullptr) {
        const float unb = M > 1 ? var * (float(M) / float(M - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
      }
    }
    stats[c] = mean;
    stats[C + c] = rsqrtf(var + eps);
  }
}

int bn_finalize(float* sums, int64_t M, int C, float eps, float momentum, float* running_mean, float* running_var, float* stats,
                int from_running, int update_running, cudaStream_t stream) {
  return int(launch_kernel(bn_finalize_kernel, dim3(unsigned((C + NT - 1) / NT)), dim3(NT), 0, stream, sums, M, C, eps, momentum,
                           running_mean, running_var, stats, from_running, update_running));
}

// Row walk shared by the element-wise kernels: thread -> fixed channel group (so per-channel constants are loaded once),
// rows r0 + rl, r0 + rl + lanes, ... inside the block's row range.
constexpr int EW_ITERS = 8;
static unsigned ew_grid(int64_t M, int C) {
  const int lanes = NT / (C >> 3) < 1 ? 1 : NT / (C >> 3);
  const int64_t rpb = int64_t(lanes) * EW_ITERS;
  return unsigned((M + rpb - 1) / rpb);
}

// z = relu?( (y - mean) * rstd * gamma + beta + residual? )
__global__ void __launch_bounds__(NT) bn_apply_kernel(const BnArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const ColLayout l = col_layout(a.C);
  if (l.rl >= l.lanes) return;
  const int c0 = l.cg << 3;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float mean, rstd;
    if (a.sums != nullptr) {
      const float inv = 1.f / float(a.M);
      mean = a.sums[c0 + i] * inv;
      const float var = fmaxf(a.sums[a.C + c0 + i] * inv - mean * mean, 0.f);
      rstd = rsqrtf(var + a.eps);
      if (blockIdx.x == 0 && l.rl == 0) {       // one thread per channel group records what backward / eval need
        a.stats_out[c0 + i] = mean;
        a.stats_out[a.C + c0 + i] = rstd;
        if (a.running_mean != nullptr) {
          const float unb = a.M > 1 ? var * (float(a.M) / float(a.M - 1)) : var;
          a.running_mean[c0 + i] = (1.f - a.momentum) * a.running_mean[c0 + i] + a.momentum * mean;
          a.running_var[c0 + i] = (1.f - a.momentum) * a.running_var[c0 + i] + a.momentum * unb;
        }
      }
    } else {
      mean = a.stats[c0 + i];
      rstd = a.stats[a.C + c0 + i];
    }
    sc[i] = rstd * a.gamma[c0 + i];
    sh[i] = a.beta[c0 + i] - mean * sc[i];
  }
  const int64_t r0 = int64_t(blockIdx.x) * l.lanes * EW_ITERS;
#pragma unroll 2
  for (int it = 0; it < EW_ITERS; ++it) {
    const int64_t m = r0 + l.rl + int64_t(it) * l.lanes;
    if (m >= a.M) break;
    float y[8], r[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(a.y + m * a.ld + c0), y);
    if (a.residual != nullptr) unpack8(*reinterpret_cast<const uint4*>(a.residual + m * a.ld_res + c0), r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = fmaf(y[i], sc[i], sh[i]);
      if (a.residual != nullptr) v += r[i];
      if (a.relu) v = fmaxf(v, 0.f);
      o[i] = v;
    }
    *reinterpret_cast<uint4*>(a.z + m * a.ld_z + c0) = pack8(o);
  }
}

int bn_apply(const BnArgs& a, cudaStream_t stream) {
  if ((a.C & 7) || a.C > 8 * NT || (a.ld & 7) || (a.ld_z & 7) || (a.ld_res & 7)) return -1;
  return int(launch_kernel(bn_apply_kernel, dim3(ew_grid(a.M, a.C)), dim3(NT), 0, stream, a));
}

// g = (dz_a + dz_b?) * (relu ? z > 0 : 1);  dbeta[c] += sum g;  dgamma[c] += sum g * xhat
__global__ void __launch_bounds__(NT) bn_bwd_reduce_kernel(const BnArgs a, int64_t rows_per_block) {
  extern __shared__ float smem[];
  pdl_wait();
  pdl_launch_dependents();
  const ColLayout l = col_layout(a.C);
  float sg[8], sx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sg[i] = 0.f; sx[i] = 0.f; }
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < a.M ? r0 + rows_per_block : a.M;
  if (l.rl < l.lanes) {
    const int c0 = l.cg << 3;
    float mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mean[i] = a.stats[c0 + i]; rstd[i] = a.stats[a.C + c0 + i]; }
    for (int64_t m = r0 + l.rl; m < r1; m += l.lanes) {
      float g[8], t[8], y[8];
      unpack8(*reinterpret_cast<const uint4*>(a.dz_a + m * a.ld_dz + c0), g);
      if (a.dz_b != nullptr) {
        unpack8(*reinterpret_cast<const uint4*>(a.dz_b + m * a.ld_dz + c0), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += t[i];
      }
      if (a.relu) {
        unpack8(*reinterpret_cast<const uint4*>(a.z + m * a.ld_z + c0), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = t[i] > 0.f ? g[i] : 0.f;
      }
      unpack8(*reinterpret_cast<const uint4*>(a.y + m * a.ld + c0), y);
#pragma unroll
      for (int i = 0; i < 8; ++i) { sg[i] += g[i]; sx[i] += g[i] * (y[i] - mean[i]) * rstd[i]; }
    }
  }
  block_col_reduce(l, sg, sx, a.dbeta, a.dgamma, smem);
}

int bn_bwd_reduce(const BnArgs& a, cudaStream_t stream) {
  if ((a.C & 7) || a.C > 8 * NT || (a.ld & 7) || (a.ld_z & 7) || (a.ld_dz & 7)) return -1;
  const int64_t rpb = rows_per_block_for(a.M, a.C);
  const unsigned grid = unsigned((a.M + rpb - 1) / rpb);
  return int(launch_kernel(bn_bwd_reduce_kernel, dim3(grid), dim3(NT), col_smem(a.C), stream, a, rpb));
}

// dy = gamma * rstd * (g - dbeta/M - xhat * dgamma/M);  optionally g_out = g (gradient of the skip connection)
__global__ void __launch_bounds__(NT) bn_bwd_apply_kernel(const BnArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const ColLayout l = col_layout(a.C);
  if (l.rl >= l.lanes) return;
  const int c0 = l.cg << 3;
  const float inv_m = 1.f / float(a.M);
  // dy = k1 * g - k2 * y + k3   with  k1 = gamma*rstd,  k2 = k1 * rstd * dgamma/M,  k3 = k2 * mean - k1 * dbeta/M
  float k1[8], k2[8], k3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mean = a.stats[c0 + i], rstd = a.stats[a.C + c0 + i];
    k1[i] = a.gamma[c0 + i] * rstd;
    k2[i] = k1[i] * rstd * a.dgamma[c0 + i] * inv_m;
    k3[i] = k2[i] * mean - k1[i] * a.dbeta[c0 + i] * inv_m;
  }
  const int64_t r0 = int64_t(blockIdx.x) * l.lanes * EW_ITERS;
#pragma unroll 2
  for (int it = 0; it < EW_ITERS; ++it) {
    const int64_t m = r0 + l.rl + int64_t(it) * l.lanes;
    if (m >= a.M) break;
    float g[8], t[8], y[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(a.dz_a + m * a.ld_dz + c0), g);
    if (a.dz_b != nullptr) {
      unpack8(*reinterpret_cast<const uint4*>(a.dz_b + m * a.ld_dz + c0), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += t[i];
    }
    if (a.relu) {
      unpack8(*reinterpret_cast<const uint4*>(a.z + m * a.ld_z + c0), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = t[i] > 0.f ? g[i] : 0.f;
    }
    unpack8(*reinterpret_cast<const uint4*>(a.y + m * a.ld + c0), y);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(k1[i], g[i], fmaf(-k2[i], y[i], k3[i]));
    *reinterpret_cast<uint4*>(a.dy + m * a.ld_dy + c0) = pack8(o);
    if (a.g_out != nullptr) *reinterpret_cast<uint4*>(a.g_out + m * a.ld_dz + c0) = pack8(g);
  }
}

int bn_bwd_apply(const BnArgs& a, cudaStream_t stream) {
  if ((a.C & 7) || a.C > 8 * NT || (a.ld & 7) || (a.ld_z & 7) || (a.ld_dz & 7) || (a.ld_dy & 7)) return -1;
  return int(launch_kernel(bn_bwd_apply_kernel, dim3(ew_grid(a.M, a.C)), dim3(NT), 0, stream, a));
}

// ------------------------------------------------------------------------------------------------ pooling
// k x k max pooling with stride s and padding p over NHWC (padding behaves like -inf); idx = argmax tap.
__global__ void __launch_bounds__(NT) maxpool_cl_fwd_kernel(const Pool2Args a) {
  pdl_wait();
  pdl_launch_dependents();
  const int c8n = a.C >> 3;
  const int64_t total = int64_t(a.B) * a.PH * a.PW * c8n;
  for (int64_t e = int64_t(blockIdx.x) * NT + threadIdx.x; e < total; e += int64_t(gridDim.x) * NT) {
    const int cg = int(e % c8n);
    int64_t t = e / c8n;
    const int pw = int(t % a.PW); t /= a.PW;
    const int ph = int(t % a.PH);
    const int64_t b = t / a.PH;
    float best[8];
    int arg[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; arg[q] = 0; }
    for (int i = 0; i < a.k; ++i)
      for (int j = 0; j < a.k; ++j) {
        const int h = ph * a.s - a.p + i, w = pw * a.s - a.p + j;
        if (h < 0 || h >= a.H || w < 0 || w >= a.W) continue;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(a.x + ((b * a.H + h) * a.W + w) * a.ld_x + (cg << 3)), f);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (f[q] > best[q]) { best[q] = f[q]; arg[q] = i * a.k + j; }
      }
    const int64_t o = ((b * a.PH + ph) * a.PW + pw);
    *reinterpret_cast<uint4*>(a.y + o * a.ld_y + (cg << 3)) = pack8(best);
    if (a.idx != nullptr) {
      uint2 pk;
      pk.x = uint32_t(arg[0]) | (uint32_t(arg[1]) << 8) | (uint32_t(arg[2]) << 16) | (uint32_t(arg[3]) << 24);
      pk.y = uint32_t(arg[4]) | (uint32_t(arg[5]) << 8) | (uint32_t(arg[6]) << 16) | (uint32_t(arg[7]) << 24);
      *reinterpret_cast<uint2*>(a.idx + o * a.C + (cg << 3)) = pk;
    }
  }
}

// gather form of the backward: every input pixel sums the gradients of the windows whose argmax it is
__global__ void __launch_bounds__(NT) maxpool_cl_bwd_kernel(const Pool2Args a) {
  pdl_wait();
  pdl_launch_dependents();
  const int c8n = a.C >> 3;
  const int64_t total = int64_t(a.B) * a.H * a.W * c8n;
  for (int64_t e = int64_t(blockIdx.x) * NT + threadIdx.x; e < total; e += int64_t(gridDim.x) * NT) {
    const int cg = int(e % c8n);
    int64_t t = e / c8n;
    const int w = int(t % a.W); t /= a.W;
    const int h = int(t % a.H);
    const int64_t b = t / a.H;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    // windows (ph, pw) with ph*s - p <= h <= ph*s - p + k - 1
    int ph_lo = (h + a.p - a.k + 1 + a.s - 1) / a.s;
    if (h + a.p - a.k + 1 < 0) ph_lo = 0;
    int pw_lo = (w + a.p - a.k + 1 + a.s - 1) / a.s;
    if (w + a.p - a.k + 1 < 0) pw_lo = 0;
    const int ph_hi = min((h + a.p) / a.s, a.PH - 1), pw_hi = min((w + a.p) / a.s, a.PW - 1);
    for (int ph = ph_lo; ph <= ph_hi; ++ph)
      for (int pw = pw_lo; pw <= pw_hi; ++pw) {
        const int tap = (h - (ph * a.s - a.p)) * a.k + (w - (pw * a.s - a.p));
        const int64_t o = ((b * a.PH + ph) * a.PW + pw);
        const uint2 pk = *reinterpret_cast<const uint2*>(a.idx + o * a.C + (cg << 3));
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(a.dy + o * a.ld_y + (cg << 3)), g);
        if (a.dy_b != nullptr) {
          float g2[8];
          unpack8(*reinterpret_cast<const uint4*>(a.dy_b + o * a.ld_y + (cg << 3)), g2);
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] += g2[q];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t id = ((q < 4 ? pk.x : pk.y) >> ((q & 3) * 8)) & 0xffu;
          if (int(id) == tap) acc[q] += g[q];
        }
      }
    *reinterpret_cast<uint4*>(a.dx + ((b * a.H + h) * a.W + w) * a.ld_x + (cg << 3)) = pack8(acc);
  }
}

int maxpool_cl(const Pool2Args& a, bool backward, cudaStream_t stream) {
  if ((a.C & 7) || (a.ld_x & 7) || (a.ld_y & 7) || a.k * a.k > 255) return -1;
  if (!backward)
    return int(launch_kernel(maxpool_cl_fwd_kernel, dim3(grid_for_total(int64_t(a.B) * a.PH * a.PW * (a.C >> 3), 16)), dim3(NT), 0,
                             stream, a));
  return int(launch_kernel(maxpool_cl_bwd_kernel, dim3(grid_for_total(int64_t(a.B) * a.H * a.W * (a.C >> 3), 16)), dim3(NT), 0, stream, a));
}

// global average pooling: x [B, HW, C] -> y [B, C];  backward: dx[b, hw, c] = dy[b, c] / HW
__global__ void __launch_bounds__(NT) avgpool_cl_kernel(__nv_bfloat16* __restrict__ x, int64_t ld_x, __nv_bfloat16* __restrict__ y,
                                                        int64_t ld_y, int B, int HW, int C, int backward) {
  pdl_wait();
  pdl_launch_dependents();
  const int c8n = C >> 3;
  const float inv = 1.f / float(HW);
  if (!backward) {
    const int64_t total = int64_t(B) * c8n;
    for (int64_t e = int64_t(blockIdx.x) * NT + threadIdx.x; e < total; e += int64_t(gridDim.x) * NT) {
      const int cg = int(e % c8n);
      const int64_t b = e / c8n;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      for (int p = 0; p < HW; ++p) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (b * HW + p) * ld_x + (cg << 3)), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += f[q];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] *= inv;
      *reinterpret_cast<uint4*>(y + b * ld_y + (cg << 3)) = pack8(acc);
    }
  } else {
    const int64_t total = int64_t(B) * HW * c8n;
    for (int64_t e = int64_t(blockIdx.x) * NT + threadIdx.x; e < total; e += int64_t(gridDim.x) * NT) {
      const int cg = int(e % c8n);
      const int64_t m = e / c8n;
      const int64_t b = m / HW;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(y + b * ld_y + (cg << 3)), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] *= inv;
      *reinterpret_cast<uint4*>(x + m * ld_x + (cg << 3)) = pack8(f);
    }
  }
}

int avgpool_cl(__nv_bfloat16* x_or_dx, int64_t ld_x, __nv_bfloat16* y_or_dy, int64_t ld_y, int B, int HW, int C, bool backward,
               cudaStream_t stream) {
  if ((C & 7) || (ld_x & 7) || (ld_y & 7)) return -1;
  const int64_t total = backward ? int64_t(B) * HW * (C >> 3) : int64_t(B) * (C >> 3);
  return int(launch_kernel(avgpool_cl_kernel, dim3(grid_for_total(total, 16)), dim3(NT), 0, stream, x_or_dx, ld_x, y_or_dy, ld_y, B, HW, C,
                           backward ? 1 : 0));
}

}  // namespace st
