// Whole-step persistent kernel for Linear/ReLU chains (see mlp_step.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace st {

constexpr int MS_THREADS = 320;          // TMA warp + MMA warp + 8 epilogue warps
constexpr int MS_STAGES = 8;
constexpr int MS_MAX_TASKS = 24;          // tasks per CTA per step (descriptor copies live in shared memory)
constexpr int MS_MAX_FLAGS = 64;          // cross-GPU flag ids; flag word = flags[dst][id * 8 + src_rank]

enum MsTaskType : int32_t { MS_END = 0, MS_GEMM = 1, MS_UPDATE = 2, MS_DECIDE = 3 };

enum MsFlags : int32_t {
  MS_A_MN = 1 << 0,        // A stored [K rows, M contiguous]
  MS_B_MN = 1 << 1,        // B stored [K rows, N contiguous]
  MS_RELU = 1 << 2,
  MS_A_WIN = 1 << 3,       // A rows offset by the mini-batch window (K-major A)
  MS_B_WIN = 1 << 4,       // B K-rows offset by the mini-batch window (MN-major B)
  MS_BIAS = 1 << 5,
  MS_MASK = 1 << 6,        // keep acc where mask[m, n] > 0 (ReLU backward)
  MS_OUT_BF16 = 1 << 7,
  MS_OUT_F32 = 1 << 8,
  MS_COLSUM = 1 << 9,      // per-tile column sums stored (not added) to a per-m-tile partial slot
  MS_LOSS_CE = 1 << 10,
  MS_LOSS_MSE = 1 << 11,
  MS_LBL_WIN = 1 << 12,    // labels / targets are the resident partition's (offset by the window)
  MS_FUSED_OPT = 1 << 13,  // one GPU: the weight-gradient tile applies the optimizer to its parameters in the epilogue
  // Chained head: the CTA that runs a tile of the last layer's dgrad (dZ_{L-1} = dZ_L W_L) recomputes the head of its 128
  // rows itself (logits, loss, dZ_L) and feeds dZ_L to the second MMA through shared memory - the step loses a whole
  // signal -> wait -> load -> store phase.  A pair of consecutive tasks on one CTA:
  MS_CHAIN_HEAD = 1 << 14, // head tile whose epilogue also writes dZ_L (bf16, 128B-swizzled K-major) into the chain buffer
  MS_QUIET = 1 << 15,      // redundant copy of a head tile: nothing is stored to global memory, nothing is signalled
  MS_CHAIN_TAIL = 1 << 16, // A operand = the chain buffer (no TMA load of A)
  MS_EARLY_A = 1 << 17,    // the A operand is the one without same-step dependencies (the input batch of the first layer): it is
                           // loaded while the task still waits for its weights, B follows once they are published
};
constexpr int MS_XCNT_BASE = 1024;       // word offset of the cross-GPU counters inside the symmetric flag block (ids * MS_CNT_STRIDE follow)
constexpr int MS_CNT_STRIDE = 32;        // one dependency counter per 128-byte line (words): polls of different counters never share a line

struct MsWait {            // kind 0: none; 1: local counter (>= epoch * need); 2: cross-GPU flag `idx` from every rank in `mask` (>= epoch);
                           // 3: cross-GPU COUNTER `idx` in this rank's symmetric block (>= (epoch - mask) * need; need = arrivals per
                           //    step from all ranks): every producer task adds 1 on every consumer rank itself - no last-arriver hop
  int32_t kind, idx, need, mask;
};

// One unit of work of a step.  256 bytes, mirrored field by field by runtime/mega_mlp.py (numpy dtype MS_TASK).
struct MsTask {
  int32_t type, flags;
  int32_t tmap_a[2], tmap_b[2];          // tensor-map table indices per buffer parity
  int32_t m0, n0, M, N;                  // tile origin, problem bounds
  int32_t num_kb, block_n;
  int32_t ld_mask, ld_out, ld_f32, ld_target;
  int64_t bias;                          // const float* (n0 is added by the kernel)
  int64_t mask[2];                       // const bf16* per parity
  int64_t out_bf16[2];                   // bf16* per parity
  int64_t out_f32;                       // float* (weight gradient, row-major, plain stores)
  int64_t colsum;                        // float* partial slot of this m-tile (n0 is added by the kernel)
  int64_t labels;                        // const int32* (CE) or const float* (MSE targets)
  int64_t loss_slot[2];                  // float* partial slot of this m-tile per parity
  float loss_scale;
  int32_t seg;                           // UPDATE: parameter segment index
  int64_t e0;                            // UPDATE: first arena element (multiple of 4)
  int32_t len, n_part;                   // UPDATE: elements (multiple of 4); gradient = sum of n_part partial rows
  int64_t part_off, part_stride;         // UPDATE: partial p >= 1 of element e lives at part_off + (p - 1) * part_stride + e
  MsWait wait[3];
  int32_t sig_cnt, sig_total, sig_flag, sig_mask;   // bump counter; when it reaches epoch * sig_total raise flag on ranks in mask
  int32_t ep_c0, ep_cols;                // epilogue window inside the tile (ep_cols = 0: the whole tile).  Two CTAs may run the same
                                         // MMA tile and drain half of its columns each: the epilogue, not the MMA, is the cost.
};
static_assert(sizeof(MsTask) == 256, "MsTask layout is mirrored in Python");

struct MsPeerTable {       // device memory; copied to shared memory by every CTA
  uint32_t* flags[8];
  float* grad[8];
  float* param[8];
  __nv_bfloat16* wbf16[8];
};

struct MsParams {
  const MsTask* tasks;     // [grid][MS_MAX_TASKS]
  const int32_t* n_tasks;  // [grid]
  const CUtensorMap* tmaps;
  const MsPeerTable* peers;
  uint32_t* counters;      // local dependency counters (monotonic)
  uint32_t* flags_local;   // = peers->flags[rank]
  int world, rank;
  float* grad_local;
  float* param_local;
  __nv_bfloat16* wbf16_local;
  float* mc_grad;          // NVLS multicast aliases (null: peer loops)
  float* mc_param;
  __nv_bfloat16* mc_wbf16;
  float* m;                // optimizer state of the owned shard, indexed by (e - shard_lo)
  float* v;
  int64_t shard_lo;
  const ParamSegment* segs;
  StepState* state;
  float* loss_hist;
  int64_t max_iters;
  volatile uint32_t* mailbox;
  int kind;                // 0 sgd, 1 adam, 2 adamw
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  int nesterov, maximize, patience;
  int decide_cnt;          // counter index carrying the early-stop decision word (epoch * 2 + stop)
  int n_mt;                // batch m-tiles = loss partial slots per parity
  int64_t loss_off;        // float offset of the loss partials in the gradient arena: [2 parities][n_mt]
  int64_t window_rows, n_windows;
  // per launch
  int64_t iter0;           // iteration number of the first step (data window, loss history index)
  int64_t step0;           // optimizer steps applied before the first step
  uint32_t epoch0;         // steps run since the counters/flags were zeroed
  int n_steps;
  long long* prof;         // optional [grid][MS_MAX_TASKS][8] globaltimer stamps of the last step
  long long* prof_kb;      // optional [grid][MS_MAX_TASKS][16][4]: per k-block B issued, A issued, stage landed (MMA thread), unused
  uint32_t* error;         // [4] device word set before a trap: code, cta, task, epoch
  int xcnt;                // 1: cross-GPU dependencies are counters (wait kind 3), tasks signal every consumer rank directly
  uint32_t* mc_flags;      // NVLS multicast alias of the symmetric flag block (null: one red per peer)
  int sys_fence_all;       // 1: every cross-GPU flag is preceded by a system-scope fence (default: only the publication's)
};

int ms_smem_bytes();
int mlp_step_launch(const MsParams& p, int grid, cudaStream_t stream);

}  // namespace st
