// Device-side mini-batch assembly (reference hot op G12: the reference samples
// with numpy on the host and fancy-indexes CPU tensors every iteration,
// distributed.py:145-148).  One kernel builds the step's inputs from the
// resident bf16 partition (or a pinned-memory staging ring): X [B,D] row-major,
// X^T [D,B] (K-major operand of the first-layer wgrad GEMM) and the targets.
// Row selection happens on the device from the iteration counter kept in
// StepState: sequential sweep, or a keyed Feistel permutation with cycle walking
// (= sampling without replacement, the reference's np.random.choice semantics).
#include "common.cuh"
#include "kernels.h"

namespace st {

ST_DEVINL uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// bijection on [0, n): 4-round Feistel over an even number of bits + cycle walking
ST_DEVINL uint64_t feistel_perm(uint64_t i, uint64_t n, uint64_t key) {
  int bits = 2;
  while ((uint64_t(1) << bits) < n) bits += 2;
  const int half = bits >> 1;
  const uint32_t mask = (uint32_t(1) << half) - 1u;
  uint64_t x = i;
  do {
    uint32_t l = uint32_t(x >> half) & mask, r = uint32_t(x) & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t f = mix32(r ^ uint32_t(key >> (16 * (round & 1))) ^ (0x9e3779b9u * (round + 1))) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (uint64_t(l) << half) | r;
  } while (x >= n);
  return x;
}

constexpr int GT_ROWS = 64, GT_COLS = 64, GT_THREADS = 256;

__global__ void __launch_bounds__(GT_THREADS) gather_batch_kernel(const GatherArgs a) {
  __shared__ __nv_bfloat16 tile[GT_ROWS][GT_COLS + 2];
  __shared__ int64_t s_src_row[GT_ROWS];
  const int b0 = blockIdx.y * GT_ROWS;
  const int c0 = blockIdx.x * GT_COLS;
  pdl_wait();
  pdl_launch_dependents();
  const int64_t iter = a.state != nullptr ? a.state->iter : 0;
  if (threadIdx.x < GT_ROWS) {
    const int b = b0 + threadIdx.x;
    int64_t src = 0;
    if (b < a.B) {
      if (a.mode == 0) src = (iter * a.B + b) % a.n_rows;
      else if (a.mode == 1) src = int64_t(feistel_perm(uint64_t(b), uint64_t(a.n_rows), a.seed + 0x632be59bd9b4e019ULL * uint64_t(iter + 1)));
      else src = (iter % a.ring) * int64_t(a.B) + b;
    }
    s_src_row[threadIdx.x] = src;
  }
  __syncthreads();
  // load: each thread moves bf16x2 pairs; 32 threads cover one 64-col row
  for (int e = threadIdx.x; e < GT_ROWS * (GT_COLS / 2); e += GT_THREADS) {
    const int r = e / (GT_COLS / 2), cp = (e % (GT_COLS / 2)) * 2;
    const int b = b0 + r, c = c0 + cp;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (b < a.B && c < a.D) {
      const __nv_bfloat16* sp = a.src + s_src_row[r] * a.ld_src + c;
      if (c + 1 < a.D) v = *reinterpret_cast<const __nv_bfloat162*>(sp);
      else v = __halves2bfloat162(sp[0], __float2bfloat16(0.f));
    }
    tile[r][cp] = v.x;
    tile[r][cp + 1] = v.y;
    if (b < a.B && c < a.ldx) {
      // row-major copy (padding columns up to ldx are written as zeros)
      __nv_bfloat16* xp = a.x + size_t(b) * a.ldx + c;
      if (c + 1 < a.ldx) *reinterpret_cast<__nv_bfloat162*>(xp) = v; else xp[0] = v.x;
    }
  }
  __syncthreads();
  if (a.xt != nullptr) {
    for (int e = threadIdx.x; e < GT_COLS * (GT_ROWS / 2); e += GT_THREADS) {
      const int c = e / (GT_ROWS / 2), rp = (e % (GT_ROWS / 2)) * 2;
      const int col = c0 + c, b = b0 + rp;
      if (col < a.D && b < a.ldxt) {
        __nv_bfloat16* tp = a.xt + size_t(col) * a.ldxt + b;
        const __nv_bfloat16 v0 = (b < a.B) ? tile[rp][c] : __float2bfloat16(0.f);
        const __nv_bfloat16 v1 = (b + 1 < a.B) ? tile[rp + 1][c] : __float2bfloat16(0.f);
        if (b + 1 < a.ldxt) *reinterpret_cast<__nv_bfloat162*>(tp) = __halves2bfloat162(v0, v1); else tp[0] = v0;
      }
    }
  }
  if (blockIdx.x == 0 && a.tgt_out != nullptr && a.tgt_src != nullptr) {
    if (a.tgt_dim == 0) {
      if (threadIdx.x < GT_ROWS && b0 + threadIdx.x < a.B)
        reinterpret_cast<int32_t*>(a.tgt_out)[b0 + threadIdx.x] = reinterpret_cast<const int32_t*>(a.tgt_src)[s_src_row[threadIdx.x]];
    } else {
      for (int e = threadIdx.x; e < GT_ROWS * a.tgt_dim; e += GT_THREADS) {
        const int r = e / a.tgt_dim, k = e % a.tgt_dim;
        if (b0 + r < a.B)
          reinterpret_cast<float*>(a.tgt_out)[size_t(b0 + r) * a.tgt_dim + k] =
              reinterpret_cast<const float*>(a.tgt_src)[s_src_row[r] * a.tgt_dim + k];
      }
    }
  }
}

int gather_batch(const GatherArgs& a, cudaStream_t stream) {
  const int64_t width = a.ldx > a.D ? a.ldx : a.D;
  dim3 grid(unsigned((width + GT_COLS - 1) / GT_COLS), unsigned((a.B + GT_ROWS - 1) / GT_ROWS));
  return int(launch_kernel(gather_batch_kernel, grid, dim3(GT_THREADS), 0, stream, a));
}

__global__ void __launch_bounds__(GT_THREADS)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t lds, int rows, int cols, __nv_bfloat16* __restrict__ dst, int64_t ldd) {
  __shared__ __nv_bfloat16 tile[GT_ROWS][GT_COLS + 2];
  const int r0 = blockIdx.y * GT_ROWS, c0 = blockIdx.x * GT_COLS;
  for (int e = threadIdx.x; e < GT_ROWS * GT_COLS; e += GT_THREADS) {
    const int r = e / GT_COLS, c = e % GT_COLS;
    tile[r][c] = (r0 + r < rows && c0 + c < cols) ? src[size_t(r0 + r) * lds + c0 + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < GT_ROWS * GT_COLS; e += GT_THREADS) {
    const int c = e / GT_ROWS, r = e % GT_ROWS;
    if (c0 + c < cols && r0 + r < ldd) dst[size_t(c0 + c) * ldd + r0 + r] = tile[r][c];
  }
}

int transpose_bf16(const __nv_bfloat16* src, int64_t lds, int rows, int cols, __nv_bfloat16* dst, int64_t ldd, cudaStream_t stream) {
  const int64_t rr = ldd > rows ? ldd : rows;
  dim3 grid(unsigned((cols + GT_COLS - 1) / GT_COLS), unsigned((rr + GT_ROWS - 1) / GT_ROWS));
  transpose_bf16_kernel<<<grid, GT_THREADS, 0, stream>>>(src, lds, rows, cols, dst, ldd);
  return int(cudaGetLastError());
}

}  // namespace st
