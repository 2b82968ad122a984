// K-STEP: one persistent kernel runs WHOLE training steps of a Linear/ReLU chain.
//
//   forward GEMMs (bias + ReLU fused) -> fused softmax-CE / MSE in the head's epilogue -> dgrad GEMMs
//   (ReLU mask + bias-gradient column sums fused) -> wgrad GEMMs (straight into the symmetric gradient
//   arena) -> cross-GPU reduce (NVLS multimem.ld_reduce or peer loads) -> Adam/AdamW/SGD on the owned
//   shard -> fp32 + bf16 publication to every replica (multimem.st) -> early-stop decision, loss history,
//   host mailbox -> next step, all inside ONE launch, for `n_steps` steps.
//
// It replaces the reference's whole hot loop (/root/reference/sparktorch/distributed.py:140-203: model(x),
// criterion, loss.item(), backward, one all_reduce per parameter, /= world, early-stop all_reduces,
// optimizer.step) and the round-1 chain of 9 dependent launches per step.
//
// Structure: a static task graph.  Every CTA owns a short list of tasks per step (tile GEMMs, update
// chunks, the decision task; built and placed by runtime/mega_mlp.py).  Dependencies are monotonic
// counters in global memory (within the GPU) and epoch flags in symmetric memory (across GPUs); nothing
// is ever reset, a wait is "counter >= epoch * need".  A CTA is warp specialised like the stand-alone
// GEMM (gemm_sm100.cu): warp 0 waits for a task's inputs and streams its operand tiles with TMA through
// an 8-stage mbarrier ring, warp 1 issues tcgen05.mma into a double-buffered TMEM accumulator, warps 2-5
// drain TMEM, run the fused epilogue, publish the tile and signal its consumers - so the epilogue of task
// t overlaps the operand loads and MMAs of task t+1.  Activations / dZ are double buffered by step
// parity; together with "forward of layer l waits for the publication of W_l" that makes every
// write-after-read hazard across steps impossible (see DESIGN.md, K-STEP).
#include "common.cuh"
#include "kernels.h"
#include "mlp_step.h"

namespace st {

namespace {

constexpr int MS_BLOCK_M = 128;
constexpr int MS_BLOCK_K = 64;
constexpr int MS_A_BYTES = MS_BLOCK_M * MS_BLOCK_K * 2;   // 16 KB
constexpr int MS_B_BYTES = 64 * MS_BLOCK_K * 2;           // 8 KB (BLOCK_N <= 64)
constexpr int MS_STAGE_BYTES = MS_A_BYTES + MS_B_BYTES;
constexpr int MS_ACC_COLS = 64;                           // TMEM columns per accumulator stage
constexpr int MS_TMEM_COLS = 2 * MS_ACC_COLS;
constexpr int MS_MISC_BYTES = 2048;                       // barriers, tmem slot, bias tile, column sums, loss, flags
constexpr int MS_SMEM_TOTAL = MS_STAGES * MS_STAGE_BYTES + MS_MISC_BYTES + int(sizeof(MsPeerTable)) +
                              MS_MAX_TASKS * int(sizeof(MsTask)) + 1024;

ST_DEVINL uint32_t ld_relaxed_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ST_DEVINL uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ST_DEVINL float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
ST_DEVINL void st_relaxed_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
ST_DEVINL void st_release_gpu_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
ST_DEVINL uint32_t atom_add_acqrel_gpu(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
ST_DEVINL void fence_acqrel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
ST_DEVINL void fence_acqrel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
ST_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// L2-only loads: everything this kernel reads that another CTA / GPU wrote during the launch bypasses L1
ST_DEVINL uint4 ld_cg_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
ST_DEVINL float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
ST_DEVINL float ld_cg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
ST_DEVINL float4 mc_ld_reduce_f4(const float* mc_ptr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_ptr) : "memory");
  return v;
}
ST_DEVINL void mc_st_f4(float* mc_ptr, const float4& g) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w) : "memory");
}
ST_DEVINL void mc_st_b64(void* mc_ptr, uint32_t lo, uint32_t hi) {   // 8 bytes of payload (four bf16), bit-exact
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc_ptr), "f"(__uint_as_float(lo)), "f"(__uint_as_float(hi)) : "memory");
}
ST_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct MsShared {
  uint64_t full_bar[MS_STAGES];
  uint64_t empty_bar[MS_STAGES];
  uint64_t tmem_full_bar[2];
  uint64_t tmem_empty_bar[2];
  uint32_t tmem_ptr;
  uint32_t stop;            // early-stop decision of the current step (update tasks)
  float loss;               // loss partial of the current tile
  float pad0;
  float bias[64];
  float colsum[64];
};
static_assert(sizeof(MsShared) <= MS_MISC_BYTES, "misc region");

// a wait that never ends is a hang of every GPU of the job: give up loudly after ~4 s
ST_DEVINL void ms_timeout(const MsParams& p, int code, int task, uint32_t epoch) {
  if (p.error != nullptr) {
    p.error[0] = uint32_t(code); p.error[1] = blockIdx.x; p.error[2] = uint32_t(task); p.error[3] = epoch;
    __threadfence_system();
  }
  __trap();
}

// mbarrier wait that cannot hang the GPU: a transaction-count or protocol bug traps instead (code 10 + role)
ST_DEVINL void ms_mbar_wait(uint64_t* bar, uint32_t parity, const MsParams& p, int code) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfff) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now; else if (now - t0 > 6000000000ull) ms_timeout(p, code, -1, 0);
    }
  }
}

// Executed by one full warp: lane 8*w + r waits for entry w / source rank r.
// `early` selects the entries to wait for: 1 = those tagged early (kind | 16: what the B operand needs), 0 = the others,
// -1 = all.
ST_DEVINL void ms_wait_warp(const MsTask* t, int ti, const MsParams& p, uint32_t epoch, int lane, int early = -1) {
  const int w = lane >> 3, r = lane & 7;
  if (w < 3) {
    int kind = t->wait[w].kind;
    if (early >= 0 && ((kind >> 4) & 1) != early) kind = 0;
    kind &= 15;
    if (kind == 1 && r == 0) {
      const uint32_t target = (epoch - uint32_t(t->wait[w].mask)) * uint32_t(t->wait[w].need);   // mask = lag in steps
      const uint32_t* c = p.counters + t->wait[w].idx;
      uint32_t spins = 0;
      unsigned long long t0 = 0;
      while (int32_t(ld_relaxed_gpu_u32(c) - target) < 0) {
        if ((++spins & 0x3ff) == 0) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) ms_timeout(p, 1, ti, epoch);
        }
      }
      fence_acqrel_gpu();
    } else if (kind == 2 && ((t->wait[w].mask >> r) & 1)) {
      const uint32_t* f = p.flags_local + t->wait[w].idx * 8 + r;
      const uint32_t target = epoch - uint32_t(t->wait[w].need);      // need = lag: 1 waits for the PREVIOUS step's flag
      uint32_t spins = 0;
      unsigned long long t0 = 0;
      while (int32_t(ld_relaxed_sys_u32(f) - target) < 0) {
        if ((++spins & 0x3ff) == 0) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) ms_timeout(p, 2, ti, epoch);
        }
      }
      if (p.world > 1) fence_acqrel_sys(); else fence_acqrel_gpu();
    }
  }
  __syncwarp();
}

// One thread, after the block-level barrier that follows the task's last store.
ST_DEVINL void ms_signal(const MsTask* t, const MsPeerTable* peers, const MsParams& p, uint32_t epoch) {
  if (t->sig_cnt < 0) return;
  if (t->sig_flag < 0) {       // nobody needs the count: release-increment, no round trip
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.counters + t->sig_cnt), "r"(1u) : "memory");
    return;
  }
  const uint32_t old = atom_add_acqrel_gpu(p.counters + t->sig_cnt, 1u);
  if (old + 1u == epoch * uint32_t(t->sig_total)) {
    if (p.world > 1) fence_acqrel_sys();
    for (int r = 0; r < p.world; ++r)
      if ((t->sig_mask >> r) & 1) st_relaxed_sys_u32(peers->flags[r] + t->sig_flag * 8 + p.rank, epoch);
  }
}

struct MsOptConsts {
  float bc1_inv, bc2_inv_sqrt;
  bool first;
};

ST_DEVINL float ms_opt_one(const MsParams& a, const MsOptConsts& oc, float p, float g, float& m, float& v) {
  if (a.maximize) g = -g;
  if (a.kind == 0) {
    if (a.weight_decay != 0.f) g += a.weight_decay * p;
    if (a.momentum != 0.f) {
      m = oc.first ? g : (a.momentum * m + (1.f - a.dampening) * g);
      g = a.nesterov ? (g + a.momentum * m) : m;
    }
    return p - a.lr * g;
  }
  if (a.kind == 2) p *= (1.f - a.lr * a.weight_decay);
  else if (a.weight_decay != 0.f) g += a.weight_decay * p;
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v) * oc.bc2_inv_sqrt + a.eps;
  return p - a.lr * oc.bc1_inv * (m / denom);
}

// sum over the data-parallel replicas of 4 consecutive gradient-arena floats starting at element e
ST_DEVINL float4 ms_reduce4(const MsParams& p, const MsPeerTable* peers, int64_t e) {
  if (p.world == 1) return ld_cg_f4(p.grad_local + e);
  if (p.mc_grad != nullptr) return mc_ld_reduce_f4(p.mc_grad + e);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < p.world; ++r) {       // canonical order: every owner adds the same numbers in the same order
    const float4 q = ld_cg_f4(peers->grad[r] + e);
    g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
  }
  return g;
}

// One 32 x CHUNK slab of the accumulator: thread = output row, registers = columns.
template <int CHUNK>
ST_DEVINL void ms_epilogue_chunk(const MsTask* t, MsShared* sh, int par, const uint32_t (&r)[32], const uint4 (&mk)[4],
                                 const __nv_bfloat16* mrow, int c0, int row, bool row_ok, int col0, int lane, int lbl_off, int y_pre) {
  const int flags = t->flags;
  const int N = t->N;
  const bool vec_ok = (col0 + CHUNK <= N);
  float acc[32];
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) {
    float v = __uint_as_float(r[j]);
    if (flags & MS_BIAS) v += sh->bias[c0 + j];
    if (flags & MS_RELU) v = fmaxf(v, 0.f);
    acc[j] = v;
  }
  if (flags & MS_MASK) {
    if (row_ok && vec_ok) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) {
        const uint4& q4 = mk[j >> 3];
        const int wi = (j & 7) >> 1;
        const uint32_t w = wi == 0 ? q4.x : (wi == 1 ? q4.y : (wi == 2 ? q4.z : q4.w));
        const float mv = __uint_as_float((j & 1) ? (w & 0xffff0000u) : (w << 16));
        acc[j] = (mv > 0.f) ? acc[j] : 0.f;
      }
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (col0 + j < N) {
          const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(mrow + col0 + j));
          acc[j] = (__uint_as_float(uint32_t(raw) << 16) > 0.f) ? acc[j] : 0.f;
        }
    }
  }
  // ---- fused loss: the thread owns the whole output row (N <= CHUNK); acc becomes dLoss/dOutput
  if (flags & MS_LOSS_CE) {
    const int y = y_pre;          // class index of this row, fetched before the accumulator was complete
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) if (j < N) mx = fmaxf(mx, acc[j]);
    float sum = 0.f, zy = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      if (j == y) zy = acc[j];
      acc[j] = (j < N) ? __expf(acc[j] - mx) : 0.f;
      sum += acc[j];
    }
    float li = row_ok ? (mx + __logf(sum) - zy) : 0.f;
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) acc[j] = (row_ok && j < N) ? (acc[j] * inv - (j == y ? 1.f : 0.f)) * t->loss_scale : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
    if (lane == 0) atomicAdd(&sh->loss, li * t->loss_scale);
  } else if (flags & MS_LOSS_MSE) {
    const float* target = reinterpret_cast<const float*>(t->labels);
    float li = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      float dlt = 0.f;
      if (row_ok && j < N) dlt = acc[j] - target[size_t(row + lbl_off) * t->ld_target + j];
      li += dlt * dlt;
      acc[j] = 2.f * dlt * t->loss_scale;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
    if (lane == 0) atomicAdd(&sh->loss, li * t->loss_scale);
  }
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) acc[j] = 0.f;
  }
  // ---- row-major bf16 (activations, dZ)
  if ((flags & MS_OUT_BF16) && row_ok) {
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(t->out_bf16[par]) + size_t(row) * t->ld_out + col0;
    if (vec_ok || (col0 + CHUNK <= t->ld_out && (flags & (MS_LOSS_CE | MS_LOSS_MSE)))) {
      // (fused loss: the padding columns up to the pitch are written as zeros)
#pragma unroll
      for (int v = 0; v < CHUNK / 8; ++v) {
        uint4 w;
        w.x = pack_bf16x2(acc[8 * v + 0], acc[8 * v + 1]);
        w.y = pack_bf16x2(acc[8 * v + 2], acc[8 * v + 3]);
        w.z = pack_bf16x2(acc[8 * v + 4], acc[8 * v + 5]);
        w.w = pack_bf16x2(acc[8 * v + 6], acc[8 * v + 7]);
        if (col0 + 8 * v < t->ld_out) reinterpret_cast<uint4*>(op)[v] = w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (col0 + j < N) op[j] = __float2bfloat16(acc[j]);
    }
  }
  // ---- row-major fp32 (weight gradient -> symmetric gradient arena, plain stores: nothing to clear, nothing to add)
  if ((flags & MS_OUT_F32) && row_ok) {
    float* op = reinterpret_cast<float*>(t->out_f32) + size_t(row) * t->ld_f32 + col0;
    if (vec_ok && (t->ld_f32 & 3) == 0) {
#pragma unroll
      for (int v = 0; v < CHUNK / 4; ++v)
        reinterpret_cast<float4*>(op)[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (col0 + j < N) op[j] = acc[j];
    }
  }
  // ---- column sums over the 32 rows of this warp (bias gradient), accumulated per tile in shared memory
  if (flags & MS_COLSUM) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (j < CHUNK) ? acc[j] : 0.f;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float send = up ? v[i] : v[i + s];
        const float keep = up ? v[i + s] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    if (lane < CHUNK) atomicAdd(&sh->colsum[c0 + lane], v[0]);
  }
}

template <int CHUNK>
ST_DEVINL void ms_epilogue_tile(const MsTask* t, MsShared* sh, int par, uint32_t tacc, int q, int lane, int lbl_off, uint64_t* tmem_empty,
                                uint64_t* tmem_full, uint32_t full_parity, const MsParams& p, long long* stamp_slot) {
  const int row = t->m0 + q * 32 + lane;
  const bool row_ok = row < t->M;
  int y_pre = 0;
  if ((t->flags & MS_LOSS_CE) && row_ok) y_pre = reinterpret_cast<const int32_t*>(t->labels)[row + lbl_off];
  const int N = t->N, n0 = t->n0, bn = t->block_n;
  const bool have_mask = (t->flags & MS_MASK) != 0;
  const __nv_bfloat16* mrow = have_mask ? reinterpret_cast<const __nv_bfloat16*>(t->mask[par]) + size_t(row_ok ? row : 0) * t->ld_mask : nullptr;
  uint4 mk[4], mk_next[4];
  auto load_mask = [&](uint4 (&dst)[4], int col0) {
    if (have_mask && row_ok && col0 + CHUNK <= N) {
#pragma unroll
      for (int v = 0; v < CHUNK / 8; ++v) dst[v] = ld_cg_u4(mrow + col0 + 8 * v);
    }
  };
  load_mask(mk_next, n0);
  // everything above (labels, first mask chunk) is in flight while the main loop of this tile still runs
  ms_mbar_wait(tmem_full, full_parity, p, 13);
  tc_fence_after();
  if (stamp_slot != nullptr) *stamp_slot = (long long)globaltimer_ns();
#pragma unroll 1
  for (int c0 = 0; c0 < bn; c0 += CHUNK) {
    const int col0 = n0 + c0;
    const bool live = col0 < N;                 // warp-uniform
    uint32_t r[32];
    if (live) {
      if constexpr (CHUNK == 32) tmem_ld_32x32(tacc + uint32_t(c0), r); else tmem_ld_32x16(tacc + uint32_t(c0), r);
#pragma unroll
      for (int v = 0; v < 4; ++v) mk[v] = mk_next[v];
      if (col0 + CHUNK < N) load_mask(mk_next, col0 + CHUNK);
      tmem_ld_wait();
    }
    if (c0 + CHUNK >= bn || col0 + CHUNK >= N) {
      // last TMEM read of this tile: hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
    if (!live) break;
    ms_epilogue_chunk<CHUNK>(t, sh, par, r, mk, mrow, c0, row, row_ok, col0, lane, lbl_off, y_pre);
    if (col0 + CHUNK >= N) break;
  }
}

}  // namespace

__global__ void __launch_bounds__(MS_THREADS, 1) mlp_step_kernel(const MsParams p_in) {
  // iter0 < 0: take the step numbers from the device-side state (what the previous launch left there) - this is what
  // lets a CUDA graph with this kernel in it be replayed step after step (input streaming, gather + step graphs).
  MsParams p = p_in;
  if (p_in.iter0 < 0) {
    p.iter0 = p_in.state->iter;
    p.step0 = p_in.state->step;
    p.epoch0 = p_in.state->pad[0];
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  MsShared* sh = reinterpret_cast<MsShared*>(smem + MS_STAGES * MS_STAGE_BYTES);
  MsPeerTable* peers = reinterpret_cast<MsPeerTable*>(smem + MS_STAGES * MS_STAGE_BYTES + MS_MISC_BYTES);
  MsTask* s_tasks = reinterpret_cast<MsTask*>(smem + MS_STAGES * MS_STAGE_BYTES + MS_MISC_BYTES + sizeof(MsPeerTable));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tasks = p.n_tasks[blockIdx.x];

  // ---- prologue: barriers, TMEM, this CTA's task descriptors and the peer pointer table -> shared memory
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < MS_STAGES; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full_bar[a], 1);
      mbar_init(&sh->tmem_empty_bar[a], 4);            // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_ptr, MS_TMEM_COLS);
    tmem_relinquish();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.tasks + size_t(blockIdx.x) * MS_MAX_TASKS);
    uint4* dst = reinterpret_cast<uint4*>(s_tasks);
    for (int i = threadIdx.x; i < n_tasks * int(sizeof(MsTask) / 16); i += MS_THREADS) dst[i] = src[i];
    const uint4* ps = reinterpret_cast<const uint4*>(p.peers);
    uint4* pd = reinterpret_cast<uint4*>(peers);
    for (int i = threadIdx.x; i < int(sizeof(MsPeerTable) / 16); i += MS_THREADS) pd[i] = ps[i];
    if (threadIdx.x < 64) { sh->bias[threadIdx.x] = 0.f; sh->colsum[threadIdx.x] = 0.f; }
    if (threadIdx.x == 0) { sh->loss = 0.f; sh->stop = 0; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_ptr;
  long long* prof = p.prof ? p.prof + size_t(blockIdx.x) * MS_MAX_TASKS * 8 : nullptr;
  long long* prof_kb = p.prof_kb ? p.prof_kb + size_t(blockIdx.x) * MS_MAX_TASKS * 64 : nullptr;

  if (warp == 0) {
    // ===================== dependency waits + TMA producer =====================
    uint32_t kbc = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const uint32_t epoch = p.epoch0 + uint32_t(s) + 1u;
      const int par = int((p.epoch0 + uint32_t(s)) & 1u);
      const int win_off = int(((p.iter0 + s) % p.n_windows) * p.window_rows);
      const bool stamp = prof != nullptr && s == p.n_steps - 1;
      for (int ti = 0; ti < n_tasks; ++ti) {
        const MsTask* t = &s_tasks[ti];
        if (t->type != MS_GEMM) continue;
        const CUtensorMap* ma = p.tmaps + t->tmap_a[par];
        const CUtensorMap* mb = p.tmaps + t->tmap_b[par];
        const bool a_mn = (t->flags & MS_A_MN) != 0, b_mn = (t->flags & MS_B_MN) != 0;
        const int a_row_off = (t->flags & MS_A_WIN) ? win_off : 0;
        const int b_k_off = (t->flags & MS_B_WIN) ? win_off : 0;
        const uint32_t tx = uint32_t(MS_A_BYTES) + (b_mn ? uint32_t(MS_B_BYTES) : uint32_t(t->block_n) * 128u);
        const int m0 = t->m0, n0 = t->n0, nkb = t->num_kb;
        const int pre = nkb < MS_STAGES ? nkb : MS_STAGES;      // stages armed before the late dependencies are met
        if (lane == 0) {
          if (stamp) prof[ti * 8 + 0] = (long long)globaltimer_ns();
          tma_prefetch_desc(ma);                   // descriptor fetch overlaps the dependency wait
          tma_prefetch_desc(mb);
        }
        // -- early: what the B operand (weights / the resident batch) needs, then B for the first `pre` k-blocks
        ms_wait_warp(t, ti, p, epoch, lane, 1);
        if (lane == 0) {
          if (stamp) prof[ti * 8 + 1] = (long long)globaltimer_ns();
          fence_proxy_async_all();
          for (int i = 0; i < pre; ++i) {
            const uint32_t kk = kbc + uint32_t(i);
            const int st_ = kk % MS_STAGES;
            ms_mbar_wait(&sh->empty_bar[st_], ((kk / MS_STAGES) & 1) ^ 1, p, 10);
            uint8_t* sb = smem + st_ * MS_STAGE_BYTES + MS_A_BYTES;
            mbar_arrive_expect_tx(&sh->full_bar[st_], tx);
            const int k0 = i * MS_BLOCK_K;
            if (b_mn) tma_load_2d(sb, mb, &sh->full_bar[st_], n0, k0 + b_k_off);
            else tma_load_2d(sb, mb, &sh->full_bar[st_], k0, n0);
            if (stamp && prof_kb && i < 16) prof_kb[ti * 64 + i * 4 + 0] = (long long)globaltimer_ns();
          }
        }
        // -- late: the A operand (what the previous phase of this step produced)
        ms_wait_warp(t, ti, p, epoch, lane, 0);
        if (lane == 0) {
          if (stamp) prof[ti * 8 + 2] = (long long)globaltimer_ns();
          fence_proxy_async_all();                 // other CTAs' generic-proxy stores -> visible to the TMA (async proxy) loads below
          for (int i = 0; i < nkb; ++i, ++kbc) {
            const int st_ = kbc % MS_STAGES;
            uint8_t* sa = smem + st_ * MS_STAGE_BYTES;
            uint8_t* sb = sa + MS_A_BYTES;
            const int k0 = i * MS_BLOCK_K;
            if (i >= pre) {
              ms_mbar_wait(&sh->empty_bar[st_], ((kbc / MS_STAGES) & 1) ^ 1, p, 10);
              mbar_arrive_expect_tx(&sh->full_bar[st_], tx);
              if (b_mn) tma_load_2d(sb, mb, &sh->full_bar[st_], n0, k0 + b_k_off);
              else tma_load_2d(sb, mb, &sh->full_bar[st_], k0, n0);
            }
            if (a_mn) {
              tma_load_2d(sa, ma, &sh->full_bar[st_], m0, k0);
              tma_load_2d(sa + MS_BLOCK_K * 128, ma, &sh->full_bar[st_], m0 + 64, k0);
            } else {
              tma_load_2d(sa, ma, &sh->full_bar[st_], k0, m0 + a_row_off);
            }
            if (stamp && prof_kb && i < 16) prof_kb[ti * 64 + i * 4 + 1] = (long long)globaltimer_ns();
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t kbc = 0, gc = 0;
      for (int s = 0; s < p.n_steps; ++s) {
        for (int ti = 0; ti < n_tasks; ++ti) {
          const MsTask* t = &s_tasks[ti];
          if (t->type != MS_GEMM) continue;
          const uint32_t as = gc & 1;
          ms_mbar_wait(&sh->tmem_empty_bar[as], ((gc >> 1) & 1) ^ 1, p, 11);
          tc_fence_after();
          const uint32_t tacc = tmem_base + as * MS_ACC_COLS;
          const bool a_mn = (t->flags & MS_A_MN) != 0, b_mn = (t->flags & MS_B_MN) != 0;
          const uint32_t bn = uint32_t(t->block_n < 16 ? 16 : t->block_n);
          const uint32_t idesc = make_idesc_bf16(MS_BLOCK_M, bn) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
          const uint64_t ka = a_mn ? 128 : 2, kb = b_mn ? 128 : 2;
          const int nkb = t->num_kb;
          for (int i = 0; i < nkb; ++i, ++kbc) {
            const int st_ = kbc % MS_STAGES;
            const uint32_t ph = (kbc / MS_STAGES) & 1;
            ms_mbar_wait(&sh->full_bar[st_], ph, p, 12);
            tc_fence_after();
            if (i == 0 && p.prof != nullptr && s == p.n_steps - 1) prof[ti * 8 + 3] = (long long)globaltimer_ns();
            if (prof_kb != nullptr && s == p.n_steps - 1 && i < 16) prof_kb[ti * 64 + i * 4 + 2] = (long long)globaltimer_ns();
            const uint32_t sa = smem_u32(smem + st_ * MS_STAGE_BYTES);
            const uint32_t sb = sa + MS_A_BYTES;
            const uint64_t da = a_mn ? make_mnmajor_sw128_desc(sa, MS_BLOCK_K * 128) : make_kmajor_sw128_desc(sa);
            const uint64_t db = b_mn ? make_mnmajor_sw128_desc(sb, MS_BLOCK_K * 128) : make_kmajor_sw128_desc(sb);
#pragma unroll
            for (int k = 0; k < MS_BLOCK_K / 16; ++k) umma_f16(tacc, da + ka * k, db + kb * k, idesc, (i | k) != 0);
            umma_commit(&sh->empty_bar[st_]);
          }
          umma_commit(&sh->tmem_full_bar[as]);
          if (p.prof != nullptr && s == p.n_steps - 1) prof[ti * 8 + 4] = (long long)globaltimer_ns();
          ++gc;
        }
      }
    }
  } else {
    // ===================== epilogue / update / decision (warps 2..5) =====================
    const int q = warp & 3;
    const int et = threadIdx.x - 64;             // 0..127
    uint32_t gc = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const uint32_t epoch = p.epoch0 + uint32_t(s) + 1u;
      const int par = int((p.epoch0 + uint32_t(s)) & 1u);
      const int64_t it = p.iter0 + s;
      const int win_off = int((it % p.n_windows) * p.window_rows);
      const bool stamp = prof != nullptr && s == p.n_steps - 1;
      for (int ti = 0; ti < n_tasks; ++ti) {
        const MsTask* t = &s_tasks[ti];
        if (t->type == MS_GEMM) {
          const uint32_t as = gc & 1;
          const int n0 = t->n0, N = t->N, bn = t->block_n;
          // The epilogue warps acquire the task's dependencies themselves (same polls as the producer warp, concurrently):
          // the bias (published by last step's update) and the ReLU mask (this step's forward) may then be fetched while
          // the main loop runs, instead of after the accumulator is complete.
          if (t->flags & (MS_BIAS | MS_MASK)) {
            if (warp == 2) ms_wait_warp(t, ti, p, epoch, lane, -1);
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
          if ((t->flags & MS_BIAS) && et < bn)
            sh->bias[et] = (n0 + et < N) ? ld_cg_f32(reinterpret_cast<const float*>(t->bias) + n0 + et) : 0.f;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const uint32_t tacc = tmem_base + (uint32_t(q * 32) << 16) + as * MS_ACC_COLS;
          const int lbl_off = (t->flags & MS_LBL_WIN) ? win_off : 0;
          long long* slot5 = (stamp && et == 0) ? &prof[ti * 8 + 5] : nullptr;
          if (bn >= 32) ms_epilogue_tile<32>(t, sh, par, tacc, q, lane, lbl_off, &sh->tmem_empty_bar[as], &sh->tmem_full_bar[as], (gc >> 1) & 1, p, slot5);
          else ms_epilogue_tile<16>(t, sh, par, tacc, q, lane, lbl_off, &sh->tmem_empty_bar[as], &sh->tmem_full_bar[as], (gc >> 1) & 1, p, slot5);
          if (t->flags & (MS_COLSUM | MS_LOSS_CE | MS_LOSS_MSE)) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if ((t->flags & MS_COLSUM) && et < bn) {
              if (n0 + et < N) reinterpret_cast<float*>(t->colsum)[n0 + et] = sh->colsum[et];
              sh->colsum[et] = 0.f;
            }
            if ((t->flags & (MS_LOSS_CE | MS_LOSS_MSE)) && et == 0) {
              *reinterpret_cast<float*>(t->loss_slot[par]) = sh->loss;
              sh->loss = 0.f;
            }
          }
          if (stamp && et == 0) prof[ti * 8 + 6] = (long long)globaltimer_ns();
          asm volatile("bar.sync 1, 128;" ::: "memory");      // every store of the tile has been issued by its thread
          if (et == 0) {
            ms_signal(t, peers, p, epoch);
            if (stamp) prof[ti * 8 + 7] = (long long)globaltimer_ns();
          }
          ++gc;
        } else if (t->type == MS_UPDATE) {
          if (stamp && et == 0) prof[ti * 8 + 0] = (long long)globaltimer_ns();
          if (warp == 2) {
            ms_wait_warp(t, ti, p, epoch, lane);
            if (lane == 0) {
              uint32_t stop = 0;
              if (p.patience > 0) {               // the early-stop decision of this step (DECIDE task on this GPU)
                const uint32_t* d = p.counters + p.decide_cnt;
                uint32_t val, spins = 0;
                unsigned long long t0 = 0;
                while (int32_t(((val = ld_relaxed_gpu_u32(d)) >> 1) - epoch) < 0) {
                  if ((++spins & 0x3ff) == 0) {
                    const unsigned long long now = globaltimer_ns();
                    if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) ms_timeout(p, 3, ti, epoch);
                  }
                }
                fence_acqrel_gpu();
                stop = val & 1u;
              }
              sh->stop = stop;
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (stamp && et == 0) prof[ti * 8 + 2] = (long long)globaltimer_ns();
          const bool skip = sh->stop != 0;
          if (!skip) {
            const ParamSegment sg = p.segs[t->seg];
            const int64_t step1 = p.step0 + s + 1;
            MsOptConsts oc{1.f, 1.f, step1 == 1};
            if (p.kind != 0) {
              oc.bc1_inv = 1.f / (1.f - powf(p.beta1, float(step1)));
              oc.bc2_inv_sqrt = rsqrtf(1.f - powf(p.beta2, float(step1)));
            }
            const float inv_w = 1.f / float(p.world);
            const bool has_m = (p.kind != 0) || (p.momentum != 0.f);
            const int n4 = t->len >> 2;
            const int64_t count = int64_t(sg.rows) * sg.cols;
            constexpr int U = 4;
            for (int base = 0; base < n4; base += 128 * U) {
              float4 g[U], pp[U], mm[U], vv[U];
              // all loads of the (up to) U iterations first: one NVLink / L2 round trip instead of U
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int i4 = base + u * 128 + et;
                if (i4 < n4) {
                  const int64_t e = t->e0 + (int64_t(i4) << 2);
                  g[u] = ms_reduce4(p, peers, e);
                  for (int pr = 1; pr < t->n_part; ++pr) {
                    const float4 x = ms_reduce4(p, peers, t->part_off + int64_t(pr - 1) * t->part_stride + e);
                    g[u].x += x.x; g[u].y += x.y; g[u].z += x.z; g[u].w += x.w;
                  }
                  pp[u] = ld_cg_f4(p.param_local + e);
                  mm[u] = has_m ? ld_cg_f4(p.m + (e - p.shard_lo)) : make_float4(0.f, 0.f, 0.f, 0.f);
                  vv[u] = (p.kind != 0) ? ld_cg_f4(p.v + (e - p.shard_lo)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int i4 = base + u * 128 + et;
                if (i4 >= n4) continue;
                const int64_t e = t->e0 + (int64_t(i4) << 2);
                float4 gq = g[u], pq = pp[u], mq = mm[u], vq = vv[u];
                gq.x *= inv_w; gq.y *= inv_w; gq.z *= inv_w; gq.w *= inv_w;
                pq.x = ms_opt_one(p, oc, pq.x, gq.x, mq.x, vq.x);
                pq.y = ms_opt_one(p, oc, pq.y, gq.y, mq.y, vq.y);
                pq.z = ms_opt_one(p, oc, pq.z, gq.z, mq.z, vq.z);
                pq.w = ms_opt_one(p, oc, pq.w, gq.w, mq.w, vq.w);
                if (has_m) *reinterpret_cast<float4*>(p.m + (e - p.shard_lo)) = mq;
                if (p.kind != 0) *reinterpret_cast<float4*>(p.v + (e - p.shard_lo)) = vq;
                // ---- publish: fp32 master copy and the bf16 GEMM operand, to every replica
                if (p.world == 1) {
                  *reinterpret_cast<float4*>(p.param_local + e) = pq;
                } else if (p.mc_param != nullptr) {
                  mc_st_f4(p.mc_param + e, pq);
                } else {
                  for (int r = 0; r < p.world; ++r) *reinterpret_cast<float4*>(peers->param[r] + e) = pq;
                }
                if (sg.w_off >= 0) {
                  const int64_t local = e - sg.offset;
                  const float pv[4] = {pq.x, pq.y, pq.z, pq.w};
                  if ((sg.cols & 3) == 0 && (sg.w_ld & 3) == 0 && (sg.w_off & 3) == 0 && local + 3 < count) {
                    const int row = int(local / sg.cols), col = int(local % sg.cols);
                    const int64_t wo = sg.w_off + int64_t(row) * sg.w_ld + col;
                    const uint32_t lo = pack_bf16x2(pv[0], pv[1]), hi = pack_bf16x2(pv[2], pv[3]);
                    if (p.world == 1) {
                      *reinterpret_cast<uint2*>(p.wbf16_local + wo) = make_uint2(lo, hi);
                    } else if (p.mc_wbf16 != nullptr) {
                      mc_st_b64(p.mc_wbf16 + wo, lo, hi);
                    } else {
                      for (int r = 0; r < p.world; ++r) *reinterpret_cast<uint2*>(peers->wbf16[r] + wo) = make_uint2(lo, hi);
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const int64_t li = local + j;
                      if (li < count) {
                        const int64_t wo = sg.w_off + int64_t(li / sg.cols) * sg.w_ld + int(li % sg.cols);
                        const __nv_bfloat16 h = __float2bfloat16(pv[j]);
                        if (p.world == 1) p.wbf16_local[wo] = h;
                        else for (int r = 0; r < p.world; ++r) peers->wbf16[r][wo] = h;
                      }
                    }
                  }
                }
              }
            }
          }
          if (stamp && et == 0) prof[ti * 8 + 6] = (long long)globaltimer_ns();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            ms_signal(t, peers, p, epoch);
            if (stamp) prof[ti * 8 + 7] = (long long)globaltimer_ns();
          }
        } else if (t->type == MS_DECIDE) {
          if (warp == 2) {
            if (stamp && lane == 0) prof[ti * 8 + 0] = (long long)globaltimer_ns();
            ms_wait_warp(t, ti, p, epoch, lane);
            if (lane == 0) {
              float loss = 0.f;
              for (int r = 0; r < p.world; ++r)
                for (int k = 0; k < p.n_mt; ++k) loss += ld_relaxed_sys_f32(peers->grad[r] + p.loss_off + int64_t(par) * p.n_mt + k);
              loss /= float(p.world);
              StepState* st = p.state;
              int stop_now = 0;
              uint32_t stop_bit = 0;
              if (st->stopped) {
                stop_bit = 1;
              } else {
                if (p.patience > 0) {
                  if (!st->es_seeded) {
                    st->es_seeded = 1; st->es_best = loss; st->es_bad = 0;
                  } else if (isnan(loss)) {
                    stop_now = 1;
                  } else if (loss < st->es_best) {
                    st->es_best = loss; st->es_bad = 0;
                  } else {
                    st->es_bad += 1;
                    if (st->es_bad >= uint32_t(p.patience)) stop_now = 1;
                  }
                }
                if (p.loss_hist != nullptr && it < p.max_iters) p.loss_hist[it] = loss;
                st->last_loss = loss;
                if (stop_now) { st->stopped = 1; stop_bit = 1; } else { st->step += 1; }
                st->iter = it + 1;
                if (p.mailbox != nullptr) {
                  p.mailbox[2] = __float_as_uint(loss);
                  p.mailbox[1] = uint32_t(it + 1);
                  p.mailbox[0] = st->stopped;
                }
              }
              st->pad[0] = epoch;                    // steps run since the counters were zeroed (advances even when stopped)
              if (p.patience > 0) st_release_gpu_u32(p.counters + p.decide_cnt, epoch * 2u + stop_bit);
              ms_signal(t, peers, p, epoch);
              if (stamp) prof[ti * 8 + 7] = (long long)globaltimer_ns();
            }
            __syncwarp();
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, MS_TMEM_COLS);
  }
}

int ms_smem_bytes() { return MS_SMEM_TOTAL; }

int mlp_step_launch(const MsParams& p, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MS_SMEM_TOTAL);
    if (e != cudaSuccess) return 1000 + int(e);
    attr_set = true;
  }
  if (grid < 1 || p.n_steps < 0) return -1;
  if (p.n_steps == 0) return 0;
  mlp_step_kernel<<<grid, MS_THREADS, MS_SMEM_TOTAL, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : 1000 + int(e);
}

}  // namespace st
