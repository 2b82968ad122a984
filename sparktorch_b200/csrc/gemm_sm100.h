// Host interface of the tcgen05 GEMM (see gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

struct StepState;                         // kernels.h

struct GemmEpilogue {
  float alpha = 1.f;
  const float* bias = nullptr;            // [N] added per column
  int relu = 0;
  const __nv_bfloat16* mask_src = nullptr;  // [M, ld_mask]: keep D[m,n] only where mask_src[m,n] > 0
  int64_t ld_mask = 0;
  __nv_bfloat16* out_bf16 = nullptr;      // [M, ld_out] row-major
  int64_t ld_out = 0;
  __nv_bfloat16* out_t_bf16 = nullptr;    // [N, ld_out_t] transposed
  int64_t ld_out_t = 0;
  float* out_f32 = nullptr;               // [M, ld_f32]
  int64_t ld_f32 = 0;
  float* out_t_f32 = nullptr;             // [N, ld_f32_t]
  int64_t ld_f32_t = 0;
  int atomic_f32 = 0;                     // fp32 outputs use red.add (needed for split-K)
  float* colsum = nullptr;                // [N] += sum_m D[m,n]
  float* colsumsq = nullptr;              // [N] += sum_m D[m,n]^2   (rows >= M contribute 0: no bias on that path)
  // fused loss on the output row (needs N <= 32, one n-tile): 0 none, 1 softmax-CE, 2 MSE.
  // acc becomes dLoss/dOut, so out_bf16 / colsum then emit dZ / db; out_f32 (if set) keeps the raw output.
  int loss_kind = 0;
  const int32_t* ce_labels = nullptr;     // [rows] class index
  const float* mse_target = nullptr;      // [rows, ld_target]
  int64_t ld_target = 0;
  float* loss_slot = nullptr;             // += loss * loss_scale
  float loss_scale = 1.f;                 // 1/M (CE) or 1/(M*N) (MSE)
  // ---- GEMM -> all-gather push (one-shot gradient exchange): every write into the local gradient arena is
  // mirrored into the receive slot this rank owns on every peer GPU (peer-mapped NVLink stores / red.add issued
  // from the epilogue, tile by tile, while other tiles are still being computed).  peer address = local + delta.
  int push_world = 0;                     // 0: disabled
  int push_me = 0;
  int push_mask = 0;                      // 1: fp32 outputs, 2: colsum, 4: loss slot
  const long long* push_delta = nullptr;  // DEVICE table [2 parities][push_world]: byte distance local arena -> my slot on that peer
                                          // (a table in global memory: dynamic indexing of kernel parameters would force the whole
                                          //  parameter struct into local memory)
  const StepState* push_state = nullptr;  // parity = iter & 1
  // optional phase timestamps (SM clock64): [cta][8] = start, prologue done, dependencies met,
  // first operand tile landed, accumulator complete, epilogue done
  long long* prof = nullptr;
};

// K-AG: the B operand (a weight matrix [N, K] of the flat fp32 parameter arena) is not read from a
// local bf16 copy but gathered inside the GEMM from the arenas of its OWNERS: the arena is sharded
// over `world` GPUs (shard_elems each), element e lives in param[e / shard_elems].  The epilogue warps
// act as a second producer during the main loop: 128-bit loads from peer memory (NVLink), fp32->bf16,
// swizzled st.shared into the operand stage, mbarrier arrive.  Optionally the gathered tile is also
// written to a local bf16 copy (wb_local) for later consumers (dgrad).
struct GemmPeerB {
  int world = 0;                          // 0: disabled
  const float* const* param = nullptr;    // DEVICE table [world] of the owners' arena base pointers
  int64_t shard_elems = 0;
  int64_t w_off = 0;                      // flat offset of W[0,0]; row pitch = K
  int64_t bias_off = -1;                  // flat offset of the bias vector (or -1)
  __nv_bfloat16* wb_local = nullptr;      // [N, ld_wb] local bf16 copy to fill (or null)
  int64_t ld_wb = 0;
};

// Mini-batch window taken on the device from StepState::iter (no gather kernel, no host in the loop).
struct GemmWindow {
  const StepState* state = nullptr;
  int64_t window_rows = 0;
  int64_t n_windows = 1;
  int flags = 0;                          // 1: offset A rows (K-major A), 2: offset B's K rows (MN-major B), 4: offset labels/targets,
                                          // 8: implicit-GEMM convolution - A is an NHWC activation tensor behind an im2col tensor map
  // flags & 8: k-block kb = filter tap kb / conv_cb (r = tap / conv_S, s = tap % conv_S), channels [64 * (kb % conv_cb), +64);
  // the tile's first output pixel m0 = (n, p, q) in an image of conv_P x conv_Q output pixels
  int conv_cb = 0, conv_S = 0, conv_P = 0, conv_Q = 0, conv_stride = 1, conv_pad = 0, conv_dil = 1;
};

// Forward convolution as an implicit GEMM: out[N*P*Q, Cout] = act(patches(x) W^T + bias), the patch matrix never exists - the
// TMA unit gathers 128 output pixels x 64 channels of one filter tap per k-block straight from x [N, H, W, C] (im2col tensor map).
// w: [Cout, R*S*C] bf16 with the reduction index ordered (r, s, c).  C % 64 == 0, square padding / stride / dilation.
int conv2d_fwd_implicit(const void* x, int N, int H, int W, int C, const void* w, int64_t ldw, int Cout, int R, int S, int stride, int pad,
                        int dil, const GemmEpilogue* ep, int block_n, cudaStream_t stream);

int gemm_pick_block_n(int M, int N, int b_mn);
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// C = act(alpha * A B^T + bias) on CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles); -2: shape not handled (gemm2cta_sm100.cu)
int gemm_bf16_2cta(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N, int K, const float* bias,
                   float alpha, int relu, cudaStream_t stream);
int gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int M, int N, int K,
              const GemmEpilogue* ep, const GemmWindow* win, int64_t a_rows_total, int64_t b_rows_total, int block_n,
              int split_k, cudaStream_t stream, const GemmPeerB* peer = nullptr);

}  // namespace st
