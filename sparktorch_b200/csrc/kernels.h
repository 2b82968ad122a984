// Host-callable launchers of the non-GEMM kernels (raw pointers + stream).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

extern bool g_use_pdl;   // launch the step kernels with programmatic dependent launch
extern int g_persistent_mode;   // GEMM: 0 never persistent, 1 auto (tiles > #SMs), 2 always
extern int g_pair_mode;         // GEMM: 1 (default) big plain products run on CTA pairs (cta_group::2), 0 never

// ---- loss.cu
int softmax_ce(const float* logits, int64_t ldz, const int32_t* labels, int B, int C, float* loss_slot,
               __nv_bfloat16* dz, int64_t ld_dz, __nv_bfloat16* dzt, int64_t ld_dzt, float* db, int pad_cols,
               cudaStream_t stream, const struct StepState* win_state = nullptr, int64_t window_rows = 0, int64_t n_windows = 1);
int mse_loss(const float* pred, int64_t ldp, const float* target, int64_t ldt, int B, int D, float* loss_slot,
             __nv_bfloat16* dz, int64_t ld_dz, __nv_bfloat16* dzt, int64_t ld_dzt, float* db, int pad_cols,
             cudaStream_t stream, const struct StepState* win_state = nullptr, int64_t window_rows = 0, int64_t n_windows = 1);

// ---- data.cu
struct StepState {            // lives in device memory, advanced by the update kernel
  uint32_t epoch;             // cross-GPU barrier epoch (monotonic)
  uint32_t blocks_done;       // last-block detection
  int64_t step;               // optimizer steps applied so far
  int64_t iter;               // iterations launched so far (sampling / staging slot index)
  uint32_t stopped;           // early stop latched
  uint32_t es_seeded;
  float es_best;
  uint32_t es_bad;
  float last_loss;
  uint32_t pad[3];
};

struct GatherArgs {
  const __nv_bfloat16* src;   // resident partition [n_rows, ld_src] or staging ring
  int64_t ld_src;
  int64_t n_rows;
  const void* tgt_src;        // int32 labels [n_rows] or fp32 targets [n_rows, tgt_dim]
  int tgt_dim;                // 0: int32 class labels
  int mode;                   // 0 sequential window, 1 random w/o replacement (Feistel), 2 staging ring slot
  int ring;                   // number of staging slots (mode 2)
  uint64_t seed;
  int B;
  int D;
  __nv_bfloat16* x;           // [B, ldx]
  int64_t ldx;
  __nv_bfloat16* xt;          // [D, ldxt]  (may be null)
  int64_t ldxt;
  void* tgt_out;              // int32 [B] or fp32 [B, tgt_dim]
  const StepState* state;
};
int gather_batch(const GatherArgs& a, cudaStream_t stream);
int transpose_bf16(const __nv_bfloat16* src, int64_t lds, int rows, int cols, __nv_bfloat16* dst, int64_t ldd, cudaStream_t stream);

// ---- comm_ar.cu
struct ParamSegment {         // one parameter tensor inside the flat arena
  int64_t offset;             // first element in the flat fp32 arena
  int32_t rows, cols;         // matrix view (bias: rows=1)
  int64_t w_off;              // element offset of the bf16 copy in the bf16 arena (-1: none)
  int32_t w_ld;
  int64_t wt_off;             // element offset of the transposed bf16 copy (-1: none)
  int32_t wt_ld;
  int32_t publish_f32;        // 1: updated fp32 value is also stored to every peer's fp32 arena
};

struct UpdateArgs {
  int world, rank;
  // per-rank symmetric buffers (device pointers valid on THIS gpu; index = peer rank)
  float* grad[8];             // gradient arena [total]
  float* param[8];            // fp32 param arena [numel]
  __nv_bfloat16* wbf16[8];    // bf16 weight arena
  uint32_t* flags[8];         // signal pads [2][8][MAX_BLOCKS]
  float* mc_grad;             // multicast (NVLS) alias of the gradient arena or null
  float* mc_param;            // multicast aliases of the parameter / bf16 arenas: updated values are published with ONE
  __nv_bfloat16* mc_wbf16;    // multimem.st through the switch instead of one store per peer (null: per-peer stores)
  float* m;                   // optimizer state (owned shard only is touched)
  float* v;
  StepState* state;
  float* loss_hist;           // [max_iters] mean loss per iteration
  int64_t max_iters;
  volatile uint32_t* host_mailbox;  // pinned: [0]=stopped, [1]=iters done, [2]=as_float(last loss)
  const ParamSegment* segs;
  int n_segs;
  int64_t numel;              // padded parameter elements (multiple of 64)
  int tail;                   // piggy-back slots after numel
  // optimizer
  int kind;                   // 0 sgd, 1 adam, 2 adamw
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  int nesterov, maximize;
  int patience;               // <=0: no early stopping
  int zero_grads;             // zero the local gradient arena after the exchange
  // one-shot "push" exchange (small arenas): peers have already written their gradients into my receive slots
  float* recv_local;          // [2 parities][world][recv_stride] on THIS gpu (slot [par][me] unused)
  int64_t recv_stride;        // floats per slot (>= numel + tail)
};
constexpr int AR_MAX_BLOCKS = 256;
constexpr int AR_FLAG_WORDS = 2 * 8 * AR_MAX_BLOCKS;
int allreduce_update(const UpdateArgs& a, cudaStream_t stream);
int allreduce_push_update(const UpdateArgs& a, cudaStream_t stream);   // one-shot: gradients were pushed by the GEMM epilogues
int publish_params(const UpdateArgs& a, cudaStream_t stream);   // (re)build bf16 copies from fp32 master, no exchange
int allreduce_sum_f32(float* const* bufs, uint32_t* const* flags, StepState* state, int world, int rank, int64_t n,
                      float* multicast, cudaStream_t stream);

// ---- conv.cu
struct Im2colArgs {
  const __nv_bfloat16* x;     // image batch addressed as x[b*sb + h*sh + w*sw + c*sc]
  int64_t sb, sh, sw, sc;
  const StepState* state;     // optional device-side batch window (sample offset)
  int64_t window_rows, n_windows;
  int B, H, W, C, kh, kw, stride, pad, OH, OW, K;   // K = C*kh*kw, columns in (c, i, j) order
  __nv_bfloat16* col;         // [B*OH*OW, ld_col]
  int64_t ld_col;
};
int im2col_nhwc(const Im2colArgs& a, cudaStream_t stream);
int flip_conv_weight(const __nv_bfloat16* w, int64_t ld_w, int Cout, int Cin, int kh, int kw, __nv_bfloat16* out, int64_t ld_o,
                     cudaStream_t stream);
struct PoolArgs {
  const __nv_bfloat16* x;     // fwd input NHWC [B, H, W, C]
  int64_t ld_x;
  __nv_bfloat16* y;           // fwd output [B, ld_y], (c, ph, pw) order
  int64_t ld_y;
  uint8_t* idx;               // [B, PH, PW, C]
  const __nv_bfloat16* dy;    // bwd input, same layout as y
  int64_t ld_dy;
  __nv_bfloat16* dx;          // bwd output NHWC
  int64_t ld_dx;
  int B, H, W, C, k, PH, PW;
  int y_nhwc;                 // pooled layout: 1 NHWC (feeds another conv), 0 (c, ph, pw) = torch flatten order
  const __nv_bfloat16* relu_x;  // bwd: forward input (post-ReLU) to mask the gradient, or null
  float* db;                  // bwd: += per-channel sums of dx (bias gradient of the producing conv), or null
  float p_drop;
  int train;
  uint64_t seed;
  const StepState* state;
};
int maxpool_dropout_fwd(const PoolArgs& a, cudaStream_t stream);
int maxpool_dropout_bwd(const PoolArgs& a, cudaStream_t stream);

// ---- norm.cu (batch norm, channels-last im2col, general pooling: residual conv nets)
struct Im2colClArgs {
  const __nv_bfloat16* x;     // NHWC [B, H, W, C], pixel pitch ld_x; C % 8 == 0
  int64_t ld_x;
  int B, H, W, C, kh, kw, stride, pad, dil, OH, OW;   // dil: input dilation (zero insertion) for strided dgrad
  __nv_bfloat16* col;         // [B*OH*OW, ld_col], column (i, j, c)
  int64_t ld_col;
};
int im2col_cl(const Im2colClArgs& a, cudaStream_t stream);
int flip_conv_weight_cl(const __nv_bfloat16* w, int64_t ld_w, int Cout, int Cin, int kh, int kw, __nv_bfloat16* out, int64_t ld_o,
                        cudaStream_t stream);
struct BnArgs {
  int64_t M;
  int C, relu;
  const __nv_bfloat16* y;     // conv output (pre-norm) [M, C]
  int64_t ld;
  __nv_bfloat16* z;           // fwd: output; bwd: forward output (ReLU mask)
  int64_t ld_z;
  const __nv_bfloat16* residual;
  int64_t ld_res;
  const float* stats;         // [2C] mean, rstd
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* dz_a;  // bwd: incoming gradient(s) wrt z
  const __nv_bfloat16* dz_b;
  int64_t ld_dz;
  __nv_bfloat16* dy;          // bwd: gradient wrt y
  int64_t ld_dy;
  __nv_bfloat16* g_out;       // bwd: masked incoming gradient (skip connection), pitch ld_dz
  float* dgamma;              // fp32 gradient arena slots
  float* dbeta;
  // bn_apply with inline finalize: statistics from column sums [2C] (sum, sum of squares) accumulated by the
  // producing GEMM's epilogue; block 0 stores mean/rstd to stats_out and updates the running statistics
  const float* sums;
  float* stats_out;
  float* running_mean;
  float* running_var;
  float eps, momentum;
};
int bn_stats(const __nv_bfloat16* y, int64_t ld, int64_t M, int C, float* sums, cudaStream_t stream);
int bn_finalize(float* sums, int64_t M, int C, float eps, float momentum, float* running_mean, float* running_var, float* stats,
                int from_running, int update_running, cudaStream_t stream);
int bn_apply(const BnArgs& a, cudaStream_t stream);
int bn_bwd_reduce(const BnArgs& a, cudaStream_t stream);
int bn_bwd_apply(const BnArgs& a, cudaStream_t stream);
struct Pool2Args {
  const __nv_bfloat16* x;     // fwd input NHWC
  __nv_bfloat16* dx;          // bwd output NHWC
  int64_t ld_x;
  __nv_bfloat16* y;           // fwd output NHWC [B, PH, PW, C]
  const __nv_bfloat16* dy;
  const __nv_bfloat16* dy_b;  // optional second incoming gradient (summed)
  int64_t ld_y;
  uint8_t* idx;               // [B, PH, PW, C] argmax tap
  int B, H, W, C, k, s, p, PH, PW;
};
int maxpool_cl(const Pool2Args& a, bool backward, cudaStream_t stream);
int avgpool_cl(__nv_bfloat16* x_or_dx, int64_t ld_x, __nv_bfloat16* y_or_dy, int64_t ld_y, int B, int HW, int C, bool backward,
               cudaStream_t stream);

// ---- comm_hw.cu (asynchronous parameter-server mode over peer memory)
constexpr int HW_MAX_STRIPES = 64;
// control block (uint32 words) in every rank's symmetric memory
constexpr int HW_CTRL_LOCKS = 0;                        // [HW_MAX_STRIPES] stripe spin locks of the owned shard
constexpr int HW_CTRL_STEPS = HW_MAX_STRIPES;           // [HW_MAX_STRIPES] optimizer steps applied per stripe
constexpr int HW_CTRL_STOP = 2 * HW_MAX_STRIPES;        // stop flag (replicated on every rank)
constexpr int HW_CTRL_PUSHES = HW_CTRL_STOP + 1;        // rank 0: total pushes
constexpr int HW_CTRL_WIN_SUM = HW_CTRL_STOP + 2;       // rank 0: float loss window sum
constexpr int HW_CTRL_WIN_CNT = HW_CTRL_STOP + 3;
constexpr int HW_CTRL_ES_LOCK = HW_CTRL_STOP + 4;
constexpr int HW_CTRL_ES_SEEDED = HW_CTRL_STOP + 5;
constexpr int HW_CTRL_ES_BEST = HW_CTRL_STOP + 6;       // float
constexpr int HW_CTRL_ES_BAD = HW_CTRL_STOP + 7;
constexpr int HW_CTRL_WORDS = HW_CTRL_STOP + 16;

struct HogwildArgs {
  int world, rank, stripes;
  float* param[8];            // owner r's fp32 arena (only shard r is authoritative)
  float* m[8];
  float* v[8];
  uint32_t* ctrl[8];
  float* local_param;         // pulled snapshot (the module's parameters are views of it)
  const float* local_grad;    // this worker's gradient arena
  const float* loss_in;       // [0] train loss, [1] validation loss of this iteration
  int use_val_loss;
  int64_t numel;
  int acquire_lock;
  int kind;
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  int nesterov, maximize;
  int patience, window_len;
  volatile uint32_t* host_mailbox;
  StepState* local_state;     // optional: iteration counter of this worker (advanced by every push)
  float* loss_hist;           // optional: [max_iters] this worker's loss per iteration
  int64_t max_iters;
  float* loss_slot_rw;        // optional: the loss slot is cleared after it was consumed (fused-loss epilogues red.add into it)
  float* local_grad_rw;       // optional: clear this worker's gradient arena after it was applied
  int dry_run;                // 1: launch only (module warm-up before graph capture)
};
int hogwild_pull(const HogwildArgs& a, cudaStream_t stream);
int hogwild_push(const HogwildArgs& a, cudaStream_t stream);

}  // namespace st
