// Fused loss forward+backward kernels (reference hot ops G4/G5/G8).
// One pass over the logits produces: the mean loss (red.add into a device slot
// that later rides in the gradient all-reduce payload), dlogits in bf16 in BOTH
// layouts the tcgen05 GEMMs want (row-major for dgrad, transposed for wgrad) and
// the bias gradient of the last layer (column sums).  The reference runs ATen
// log_softmax + nll_loss fwd/bwd (4 kernels) after a Python exception on the
// first attempt with float labels (quirk Q3), plus a separate bias reduction.
#include "common.cuh"
#include "kernels.h"

namespace st {

constexpr int LOSS_THREADS = 256;
constexpr int MAX_SMEM_CLASSES = 2048;

// logits [B, ldz] fp32; labels [B] int32; dz [B, ld_dz] bf16; dzt [C, ld_dzt] bf16; db [C] fp32; loss_slot fp32
// grad_scale multiplies the gradient (1/B for mean reduction); `train`=0 -> loss only.
__global__ void __launch_bounds__(LOSS_THREADS)
softmax_ce_kernel(const float* __restrict__ logits, int64_t ldz, const int32_t* __restrict__ labels, int B, int C,
                  float* __restrict__ loss_slot, __nv_bfloat16* __restrict__ dz, int64_t ld_dz,
                  __nv_bfloat16* __restrict__ dzt, int64_t ld_dzt, float* __restrict__ db, int pad_cols,
                  const StepState* __restrict__ win_state, int64_t window_rows, int64_t n_windows) {
  __shared__ float s_db[MAX_SMEM_CLASSES];
  __shared__ float s_loss;
  const bool want_db = db != nullptr && C <= MAX_SMEM_CLASSES;
  pdl_wait();
  pdl_launch_dependents();
  // labels of the resident partition: the batch is the window of the current iteration (same rule as GemmWindow)
  if (win_state != nullptr) labels += (win_state->iter % n_windows) * window_rows;
  for (int c = threadIdx.x; c < C && want_db; c += blockDim.x) s_db[c] = 0.f;
  if (threadIdx.x == 0) s_loss = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_block;
  const float inv_b = 1.f / float(B);
  float loss_acc = 0.f;
  for (int row = gwarp; row < B; row += nwarps) {
    const float* z = logits + size_t(row) * ldz;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += __expf(z[c] - mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int y = labels[row];
    const float lse = mx + __logf(sum);
    if (lane == 0) loss_acc += (lse - z[y]);
    if (dz != nullptr || dzt != nullptr) {
      const float inv_sum = 1.f / sum;
      for (int c = lane; c < C; c += 32) {
        const float g = (__expf(z[c] - mx) * inv_sum - (c == y ? 1.f : 0.f)) * inv_b;
        if (dz != nullptr) dz[size_t(row) * ld_dz + c] = __float2bfloat16(g);
        if (dzt != nullptr) dzt[size_t(c) * ld_dzt + row] = __float2bfloat16(g);
        if (want_db) atomicAdd(&s_db[c], g);
        else if (db != nullptr) red_add_f32(db + c, g);
      }
      if (dz != nullptr)
        for (int c = C + lane; c < pad_cols; c += 32) dz[size_t(row) * ld_dz + c] = __float2bfloat16(0.f);
    }
  }
  if (lane == 0) atomicAdd(&s_loss, loss_acc);
  __syncthreads();
  if (want_db)
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      if (s_db[c] != 0.f) red_add_f32(db + c, s_db[c]);
  if (threadIdx.x == 0) red_add_f32(loss_slot, s_loss * inv_b);
}

// pred [B, ldp] fp32, target [B, ldt] fp32, D outputs.  loss = mean((p-t)^2) over B*D.
__global__ void __launch_bounds__(LOSS_THREADS)
mse_kernel(const float* __restrict__ pred, int64_t ldp, const float* __restrict__ target, int64_t ldt, int B, int D,
           float* __restrict__ loss_slot, __nv_bfloat16* __restrict__ dz, int64_t ld_dz, __nv_bfloat16* __restrict__ dzt,
           int64_t ld_dzt, float* __restrict__ db, int pad_cols, const StepState* __restrict__ win_state, int64_t window_rows,
           int64_t n_windows) {
  __shared__ float s_db[MAX_SMEM_CLASSES];
  __shared__ float s_loss;
  const bool want_db = db != nullptr && D <= MAX_SMEM_CLASSES;
  pdl_wait();
  pdl_launch_dependents();
  if (win_state != nullptr) target += (win_state->iter % n_windows) * window_rows * ldt;
  for (int c = threadIdx.x; c < D && want_db; c += blockDim.x) s_db[c] = 0.f;
  if (threadIdx.x == 0) s_loss = 0.f;
  __syncthreads();
  const float inv = 1.f / (float(B) * float(D));
  float acc = 0.f;
  const int64_t total = int64_t(B) * pad_cols;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int row = int(i / pad_cols), c = int(i % pad_cols);
    if (c < D) {
      const float d = pred[size_t(row) * ldp + c] - target[size_t(row) * ldt + c];
      acc += d * d;
      const float g = 2.f * d * inv;
      if (dz != nullptr) dz[size_t(row) * ld_dz + c] = __float2bfloat16(g);
      if (dzt != nullptr) dzt[size_t(c) * ld_dzt + row] = __float2bfloat16(g);
      if (want_db) atomicAdd(&s_db[c], g);
      else if (db != nullptr) red_add_f32(db + c, g);
    } else if (dz != nullptr) {
      dz[size_t(row) * ld_dz + c] = __float2bfloat16(0.f);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_loss, acc);
  __syncthreads();
  if (want_db)
    for (int c = threadIdx.x; c < D; c += blockDim.x)
      if (s_db[c] != 0.f) red_add_f32(db + c, s_db[c]);
  if (threadIdx.x == 0) red_add_f32(loss_slot, s_loss * inv);
}

static int grid_for(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > 148 * 4) g = 148 * 4;
  return int(g);
}

int softmax_ce(const float* logits, int64_t ldz, const int32_t* labels, int B, int C, float* loss_slot,
               __nv_bfloat16* dz, int64_t ld_dz, __nv_bfloat16* dzt, int64_t ld_dzt, float* db, int pad_cols,
               cudaStream_t stream, const StepState* win_state, int64_t window_rows, int64_t n_windows) {
  if (n_windows < 1) n_windows = 1;
  return int(launch_kernel(softmax_ce_kernel, dim3(grid_for(B, LOSS_THREADS / 32)), dim3(LOSS_THREADS), 0, stream, logits, ldz,
                           labels, B, C, loss_slot, dz, ld_dz, dzt, ld_dzt, db, pad_cols, win_state, window_rows, n_windows));
}

int mse_loss(const float* pred, int64_t ldp, const float* target, int64_t ldt, int B, int D, float* loss_slot,
             __nv_bfloat16* dz, int64_t ld_dz, __nv_bfloat16* dzt, int64_t ld_dzt, float* db, int pad_cols,
             cudaStream_t stream, const StepState* win_state, int64_t window_rows, int64_t n_windows) {
  if (n_windows < 1) n_windows = 1;
  return int(launch_kernel(mse_kernel, dim3(grid_for(int64_t(B) * pad_cols, LOSS_THREADS)), dim3(LOSS_THREADS), 0, stream, pred,
                           ldp, target, ldt, B, D, loss_slot, dz, ld_dz, dzt, ld_dzt, db, pad_cols, win_state, window_rows, n_windows));
}

}  // namespace st
