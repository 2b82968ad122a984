// Convolution support kernels (reference hot ops G6/G7).
//
// Conv2d forward / backward run on the tcgen05 GEMM (gemm_sm100.cu) as implicit
// GEMMs over channels-last (NHWC) activations:
//
//   forward : Y[M, Cout]   = col(X)[M, Cin*kh*kw] * W[Cout, Cin*kh*kw]^T        (+bias, +ReLU fused)
//   wgrad   : dW[Cout, K]  = dY^T * col(X)            (both operands MN-major, split-K red.add)
//   dgrad   : dX[M', Cin]  = col_pad(dY)[M', Cout*kh*kw] * Wflip[Cin, Cout*kh*kw]^T
//                            (a "full" correlation with flipped taps: no scatter / atomics; the ReLU
//                             mask of the previous activation and its bias-gradient column sums are
//                             fused in the GEMM epilogue)
//
// This file provides the operand builders around those GEMMs:
//   im2col_nhwc      patch matrix in (c, i, j) column order = torch's weight layout, so the bf16
//                    weight copy and the fp32 weight gradient need no permutation.
//   flip_conv_weight Wflip[ci, (co, kh-1-i, kw-1-j)] = W[co, ci, i, j]
//   maxpool(+dropout2d) forward/backward: pooled values are written directly in torch's
//                    flatten order (c, h, w) so the following Linear layer sees [B, C*PH*PW].
// The reference runs cuDNN conv fwd / bwd-data / bwd-filter plus ATen pooling, Philox dropout
// and flatten kernels for the same work (examples/cnn_network.py:16-23).
#include "common.cuh"
#include "kernels.h"

namespace st {

constexpr int CV_THREADS = 256;

// col[m, c*kh*kw + i*kw + j] = x[b, oh*stride - pad + i, ow*stride - pad + j, c]   (0 outside)
// x is addressed with explicit element strides (sb, sh, sw, sc), so NHWC activations, NCHW-flat
// Spark rows and pooled (c,h,w)-ordered maps are all valid sources.  With a StepState the sample
// index is offset by the device-side mini-batch window (see GemmWindow).
// Thread <-> column k (so its (c, i, j) decomposition is computed once and global writes of a warp are
// contiguous along k); a block walks IM2COL_ROWS output pixels.  Source reads of neighbouring pixels overlap
// heavily and are served by L1.
constexpr int IM2COL_ROWS = 64;
__global__ void __launch_bounds__(CV_THREADS) im2col_nhwc_kernel(const Im2colArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t b_off = a.state != nullptr ? (a.state->iter % a.n_windows) * a.window_rows : 0;
  const int64_t M = int64_t(a.B) * a.OH * a.OW;
  const int kk = a.kh * a.kw;
  const int64_t m_begin = int64_t(blockIdx.x) * IM2COL_ROWS;
  const int64_t m_end = m_begin + IM2COL_ROWS < M ? m_begin + IM2COL_ROWS : M;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  // 2-D thread layout: tk threads along k (power of two covering the row pitch, <= 256), the rest along rows
  int tk = 1;
  while (tk < a.ld_col && tk < CV_THREADS) tk <<= 1;
  const int tr = CV_THREADS / tk;
  const int tx = threadIdx.x % tk, ty = threadIdx.x / tk;
  for (int k = tx; k < a.ld_col; k += tk) {
    const bool live = k < a.K;
    const int c = live ? k / kk : 0, r = live ? k - c * kk : 0;
    const int i = r / a.kw, j = r - i * a.kw;
    const int64_t coff = int64_t(c) * a.sc;
    for (int64_t m = m_begin + ty; m < m_end; m += tr) {
      __nv_bfloat16 v = zero;
      if (live) {
        const int ow = int(m % a.OW);
        const int64_t t = m / a.OW;
        const int oh = int(t % a.OH);
        const int64_t b = t / a.OH;
        const int h = oh * a.stride - a.pad + i, w = ow * a.stride - a.pad + j;
        if (h >= 0 && h < a.H && w >= 0 && w < a.W) v = a.x[(b + b_off) * a.sb + int64_t(h) * a.sh + int64_t(w) * a.sw + coff];
      }
      a.col[m * a.ld_col + k] = v;
    }
  }
}

int im2col_nhwc(const Im2colArgs& a, cudaStream_t stream) {
  const int64_t M = int64_t(a.B) * a.OH * a.OW;
  const int64_t g = (M + IM2COL_ROWS - 1) / IM2COL_ROWS;
  return int(launch_kernel(im2col_nhwc_kernel, dim3(unsigned(g < 1 ? 1 : g)), dim3(CV_THREADS), 0, stream, a));
}

// w: [Cout, ld_w] bf16 holding [Cout, Cin, kh, kw]; out: [Cin, ld_o] holding [Cin, Cout, kh, kw] with flipped taps
__global__ void __launch_bounds__(CV_THREADS)
flip_conv_weight_kernel(const __nv_bfloat16* __restrict__ w, int64_t ld_w, int Cout, int Cin, int kh, int kw,
                        __nv_bfloat16* __restrict__ out, int64_t ld_o) {
  pdl_wait();
  pdl_launch_dependents();
  const int kk = kh * kw;
  const int total = Cin * Cout * kk;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int ci = e / (Cout * kk);
    const int r = e - ci * Cout * kk;
    const int co = r / kk, t = r - co * kk;
    const int i = t / kw, j = t - i * kw;
    out[int64_t(ci) * ld_o + co * kk + i * kw + j] = w[int64_t(co) * ld_w + ci * kk + (kh - 1 - i) * kw + (kw - 1 - j)];
  }
}

int flip_conv_weight(const __nv_bfloat16* w, int64_t ld_w, int Cout, int Cin, int kh, int kw, __nv_bfloat16* out, int64_t ld_o,
                     cudaStream_t stream) {
  const int total = Cin * Cout * kh * kw;
  int g = (total + CV_THREADS - 1) / CV_THREADS;
  if (g > 148 * 4) g = 148 * 4;
  return int(launch_kernel(flip_conv_weight_kernel, dim3(unsigned(g)), dim3(CV_THREADS), 0, stream, w, ld_w, Cout, Cin, kh, kw, out, ld_o));
}

ST_DEVINL uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// x: NHWC [B, H, W, C] (pitch ld_x per pixel).  k x k max pooling with stride k (floor), then dropout2d
// (whole channels of a sample are zeroed with probability p, survivors scaled by 1/(1-p)).
// y: [B, ld_y] with element (c, ph, pw) at c*PH*PW + ph*PW + pw (torch flatten order).
// idx: [B, PH, PW, C] uint8 = argmax position inside the window, 255 = dropped channel.
__global__ void __launch_bounds__(CV_THREADS) maxpool_dropout_fwd_kernel(const PoolArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = int64_t(a.B) * a.PH * a.PW * a.C;
  const uint32_t iter = a.state != nullptr ? uint32_t(a.state->iter) : 0u;
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint32_t thresh = a.p_drop > 0.f ? uint32_t(double(a.p_drop) * 4294967296.0) : 0u;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(e % a.C);
    int64_t t = e / a.C;
    const int pw = int(t % a.PW); t /= a.PW;
    const int ph = int(t % a.PH);
    const int64_t b = t / a.PH;
    float best = -INFINITY;
    int arg = 0;
    for (int i = 0; i < a.k; ++i)
      for (int j = 0; j < a.k; ++j) {
        const float v = __bfloat162float(a.x[((b * a.H + ph * a.k + i) * a.W + pw * a.k + j) * a.ld_x + c]);
        if (v > best) { best = v; arg = i * a.k + j; }
      }
    bool dropped = false;
    if (a.train && a.p_drop > 0.f) dropped = hash3(uint32_t(a.seed) + iter * 0x632BE5ABu, uint32_t(b), uint32_t(c)) < thresh;
    const float outv = dropped ? 0.f : best * (a.train ? keep_scale : 1.f);
    // NHWC output: one row per pooled pixel (pitch ld_y); flatten output: one row per sample (pitch ld_y)
    const int64_t yo = a.y_nhwc ? ((b * a.PH + ph) * a.PW + pw) * a.ld_y + c
                                : b * a.ld_y + int64_t(c) * a.PH * a.PW + ph * a.PW + pw;
    a.y[yo] = __float2bfloat16(outv);
    if (a.idx != nullptr) a.idx[e] = dropped ? uint8_t(255) : uint8_t(arg);
  }
}

// dy: [B, ld_dy] in (c, ph, pw) order; dx: NHWC [B, H, W, C] (pitch ld_dx) – fully written for the pooled region.
// With relu_x (the forward input = post-ReLU activation of the producing conv) the gradient is masked by
// (x[argmax] > 0); with db the per-channel sums of the scattered gradient (= that conv's bias gradient)
// are accumulated (shared-memory partials, one red.add per channel per block).
__global__ void __launch_bounds__(CV_THREADS) maxpool_dropout_bwd_kernel(const PoolArgs a) {
  __shared__ float s_db[1024];
  pdl_wait();
  pdl_launch_dependents();
  const bool want_db = a.db != nullptr && a.C <= 1024;
  for (int c = threadIdx.x; c < a.C && want_db; c += blockDim.x) s_db[c] = 0.f;
  __syncthreads();
  const int64_t total = int64_t(a.B) * a.PH * a.PW * a.C;
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(e % a.C);
    int64_t t = e / a.C;
    const int pw = int(t % a.PW); t /= a.PW;
    const int ph = int(t % a.PH);
    const int64_t b = t / a.PH;
    const uint8_t id = a.idx[e];
    const int64_t yo = a.y_nhwc ? ((b * a.PH + ph) * a.PW + pw) * a.ld_dy + c
                                : b * a.ld_dy + int64_t(c) * a.PH * a.PW + ph * a.PW + pw;
    float g = (id == 255) ? 0.f : __bfloat162float(a.dy[yo]) * keep_scale;
    if (a.relu_x != nullptr && id != 255) {
      const int i = id / a.k, j = id % a.k;
      if (!(__bfloat162float(a.relu_x[((b * a.H + ph * a.k + i) * a.W + pw * a.k + j) * a.ld_dx + c]) > 0.f)) g = 0.f;
    }
    for (int i = 0; i < a.k; ++i)
      for (int j = 0; j < a.k; ++j)
        a.dx[((b * a.H + ph * a.k + i) * a.W + pw * a.k + j) * a.ld_dx + c] = __float2bfloat16((i * a.k + j == id) ? g : 0.f);
    if (want_db && g != 0.f) atomicAdd(&s_db[c], g);
    else if (a.db != nullptr && !want_db && g != 0.f) red_add_f32(a.db + c, g);
  }
  __syncthreads();
  if (want_db)
    for (int c = threadIdx.x; c < a.C; c += blockDim.x)
      if (s_db[c] != 0.f) red_add_f32(a.db + c, s_db[c]);
}

static unsigned pool_grid(const PoolArgs& a) {
  const int64_t total = int64_t(a.B) * a.PH * a.PW * a.C;
  int64_t g = (total + CV_THREADS - 1) / CV_THREADS;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return unsigned(g);
}

int maxpool_dropout_fwd(const PoolArgs& a, cudaStream_t stream) {
  return int(launch_kernel(maxpool_dropout_fwd_kernel, dim3(pool_grid(a)), dim3(CV_THREADS), 0, stream, a));
}
int maxpool_dropout_bwd(const PoolArgs& a, cudaStream_t stream) {
  return int(launch_kernel(maxpool_dropout_bwd_kernel, dim3(pool_grid(a)), dim3(CV_THREADS), 0, stream, a));
}

}  // namespace st
