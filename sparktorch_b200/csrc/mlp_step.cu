// K-STEP: one persistent kernel runs WHOLE training steps of a Linear/ReLU chain.
//
//   forward GEMMs (bias + ReLU fused) -> fused softmax-CE / MSE in the head's epilogue -> dgrad GEMMs
//   (ReLU mask + bias-gradient column sums fused) -> wgrad GEMMs -> [1 GPU: Adam/SGD fused into the wgrad
//   epilogue | N GPUs: symmetric gradient arena -> NVLS multimem.ld_reduce / peer loads -> Adam/SGD on the
//   owned shard -> fp32 + bf16 publication to every replica with multimem.st] -> early-stop decision, loss
//   history, host mailbox -> next step, all inside ONE launch, for `n_steps` steps.
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
// an 8-stage mbarrier ring, warp 1 issues tcgen05.mma into a double-buffered TMEM accumulator, warps 2-9
// drain TMEM, run the fused epilogue, publish the tile and signal its consumers - so the epilogue of task
// t overlaps the operand loads and MMAs of task t+1.  Activations / dZ are double buffered by step
// parity; together with "forward of layer l waits for the publication of W_l" that makes every
// write-after-read hazard across steps impossible (see DESIGN.md, K-STEP).
//
// Two things the first version of this file taught (profiles/kstep_v1_timeline_n1.md): the async
// instructions (UTCHMMA, UTMALDG, UTCBAR) must be issued from warp-UNIFORM code (whole warp runs the loop,
// `elect.sync` picks the issuing lane) - from a `lane == 0` branch every one of them is wrapped in an
// R2UR waterfall loop; and the code must stay small - 300 KB of fully unrolled epilogues missed the
// instruction cache on every task.  Hence the rolled 8-column epilogue below (and the few hot shapes as lean templates).
//
// Later cuts, each kept only after the kernel's own timeline showed the gain (profiles/round2_summary.md):
//   chained head   the CTAs of the last layer's dgrad tiles recompute the head of their rows (logits, loss, dZ_L) and feed
//                  dZ_L to the dgrad MMA through shared memory (MS_CHAIN_HEAD / MS_QUIET / MS_CHAIN_TAIL): one phase less;
//   windows        a backward tile is run by two CTAs that drain half of its columns each (ep_c0 / ep_cols);
//   early A        first-layer tiles load the batch before their weights are published (MS_EARLY_A);
//   fences         consumers acquire at gpu scope; only the publication's flag is preceded by a system-scope fence.
#include <cstddef>

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
constexpr int MS_EPI_THREADS = MS_THREADS - 64;          // warps 2.. : two groups of four warps (one per TMEM lane quadrant)
constexpr int MS_CHAIN_BYTES = MS_BLOCK_M * 128;          // dZ_L of a chained head: 128 rows x one 128-byte swizzle row (K <= 64)
constexpr int MS_CHAIN_OFF = MS_STAGES * MS_STAGE_BYTES;  // 1024-byte aligned (the stages are)
constexpr int MS_MISC_OFF = MS_CHAIN_OFF + MS_CHAIN_BYTES;
constexpr int MS_SMEM_TOTAL = MS_MISC_OFF + MS_MISC_BYTES + int(sizeof(MsPeerTable)) +
                              MS_MAX_TASKS * int(sizeof(MsTask)) + 1024;
static_assert((MS_STAGES & (MS_STAGES - 1)) == 0, "ring index uses a mask");

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
// 32 lanes x 8 consecutive fp32 columns of the accumulator
ST_DEVINL void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
// warp-uniform broadcast (lane 0's value): lets the compiler keep loop state of the issuing warps in uniform registers
ST_DEVINL int ubc(int v) { return __shfl_sync(0xffffffffu, v, 0); }

struct MsShared {
  uint64_t full_bar[MS_STAGES];
  uint64_t empty_bar[MS_STAGES];
  uint64_t tmem_full_bar[2];
  uint64_t tmem_empty_bar[2];
  uint64_t chain_bar;       // "dZ_L is in the chain buffer": one arrive per epilogue warp of the first group
  uint32_t tmem_ptr;
  uint32_t stop;            // early-stop decision of the current step (update tasks)
  uint32_t pad_[4];         // (keeps bias / colsum 16-byte aligned for the vector loads of the lean epilogues)
  float loss_w[4];          // loss partial of the current tile, one slot per epilogue warp of the first group
  float bias[64];
  float colsum[4][64];      // column sums of the current tile per TMEM lane quadrant (summed in a fixed order: run-to-run bit-exact)
};
static_assert(sizeof(MsShared) <= MS_MISC_BYTES, "misc region");
static_assert(offsetof(MsShared, bias) % 16 == 0 && offsetof(MsShared, colsum) % 16 == 0, "vector loads");

// a wait that never ends is a hang of every GPU of the job: give up loudly
__device__ __noinline__ void ms_timeout(uint32_t* error, int code, int task, uint32_t epoch) {
  if (error != nullptr) {
    error[0] = uint32_t(code); error[1] = blockIdx.x; error[2] = uint32_t(task); error[3] = epoch;
    __threadfence_system();
  }
  __trap();
}

__device__ __noinline__ void ms_mbar_wait_slow(uint64_t* bar, uint32_t parity, uint32_t* error, int code) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfff) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now; else if (now - t0 > 6000000000ull) ms_timeout(error, code, -1, 0);
    }
  }
}
// mbarrier wait that cannot hang the GPU: a transaction-count or protocol bug traps instead (code 10 + role)
ST_DEVINL void ms_mbar_wait(uint64_t* bar, uint32_t parity, const MsParams& p, int code) {
  if (!mbar_try_wait(bar, parity)) ms_mbar_wait_slow(bar, parity, p.error, code);
}

// Executed by one full warp: lane 8*w + r waits for entry w / source rank r.
// `early` selects the entries to wait for: 1 = those tagged early (kind | 16: what the B operand needs), 0 = the others,
// -1 = all.  The lanes poll CONVERGENTLY (one loop for the whole warp, every lane its own word) and the warp executes ONE
// acquire fence at the end: per-lane spin loops would run one after the other, and with them up to 8 system-scope fences.
ST_DEVINL void ms_wait_warp(const MsTask* t, int ti, const MsParams& p, uint32_t epoch, int lane, int early) {
  const int w = lane >> 3, r = lane & 7;
  const uint32_t* ptr = nullptr;
  uint32_t target = 0;
  bool sys = false;
  if (w < 3) {
    int kind = t->wait[w].kind;
    if (early >= 0 && ((kind >> 4) & 1) != early) kind = 0;
    kind &= 15;
    if (kind == 1 && r == 0) {
      target = (epoch - uint32_t(t->wait[w].mask)) * uint32_t(t->wait[w].need);   // mask = lag in steps
      ptr = p.counters + t->wait[w].idx * MS_CNT_STRIDE;
    } else if (kind == 2 && ((t->wait[w].mask >> r) & 1)) {
      target = epoch - uint32_t(t->wait[w].need);      // need = lag: 1 waits for the PREVIOUS step's flag
      ptr = p.flags_local + t->wait[w].idx * 8 + r;
      sys = true;
    } else if (kind == 3 && r == 0) {
      target = (epoch - uint32_t(t->wait[w].mask)) * uint32_t(t->wait[w].need);
      ptr = p.flags_local + MS_XCNT_BASE + t->wait[w].idx * MS_CNT_STRIDE;
      sys = true;
    }
  }
  const bool any = __any_sync(0xffffffffu, ptr != nullptr);
  if (!any) return;
  const bool any_sys = __any_sync(0xffffffffu, sys) && p.world > 1;
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    bool ok = true;
    if (ptr != nullptr) ok = int32_t((sys ? ld_relaxed_sys_u32(ptr) : ld_relaxed_gpu_u32(ptr)) - target) >= 0;
    if (__all_sync(0xffffffffu, ok)) break;
    if ((++spins & 0x3ff) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) ms_timeout(p.error, sys ? 2 : 1, ti, epoch);
    }
  }
  // Consumer side: everything a task reads after this wait lives in THIS GPU's memory (activations, the local replicas of
  // weights / biases that the owners published with remote stores, the flags themselves) or is fetched with multimem.ld_reduce,
  // which does not go through this SM's caches - the coherence point for all of it is the local L2, so a gpu-scope acquire
  // is what is needed here.  The system-scope fence is on the PRODUCER side (ms_signal), before a flag leaves the GPU.
  // (A system-scope fence here sat on the critical path of every tile at world > 1: +1.5-2 us per phase.)
  (void)any_sys;
  fence_acqrel_gpu();
}

// One thread, after the block-level barrier that follows the task's last store.
ST_DEVINL void ms_signal(const MsTask* t, const MsPeerTable* peers, const MsParams& p, uint32_t epoch) {
  if (t->sig_cnt < 0) return;
  if (t->sig_flag < 0) {       // nobody needs the count: release-increment, no round trip
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.counters + t->sig_cnt * MS_CNT_STRIDE), "r"(1u) : "memory");
    return;
  }
  if (p.xcnt) {
    // every task tells every consumer rank itself: one system-scope fence (this CTA's stores, made visible to this thread by the
    // block barrier, are ordered before the increments), then relaxed adds - through the switch's multicast when all ranks listen
    if (p.world > 1) fence_acqrel_sys(); else fence_acqrel_gpu();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p.counters + t->sig_cnt * MS_CNT_STRIDE), "r"(1u) : "memory");
    const int off = MS_XCNT_BASE + t->sig_flag * MS_CNT_STRIDE;
    if (p.mc_flags != nullptr && t->sig_mask == (1 << p.world) - 1) {
      asm volatile("multimem.red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p.mc_flags + off), "r"(1u) : "memory");
    } else {
      for (int r = 0; r < p.world; ++r)
        if ((t->sig_mask >> r) & 1) asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(peers->flags[r] + off), "r"(1u) : "memory");
    }
    return;
  }
  const uint32_t old = atom_add_acqrel_gpu(p.counters + t->sig_cnt * MS_CNT_STRIDE, 1u);
  if (old + 1u == epoch * uint32_t(t->sig_total)) {
    // What a flag announces to another GPU is either data in THIS GPU's memory (activation gradients, weight-gradient tiles, loss
    // partials: peers fetch them through this GPU's L2, where the acq_rel atomic above has already ordered every producer's
    // release) or - update tasks only - data this rank stored into the PEERS' memories (the publication): only that needs the
    // system-scope fence, which waits for the remote writes to be acknowledged (2-3 us, measured with SPARKTORCH_MEGA_XCNT=1
    // where every task paid it: 41.7 -> 49.6 us per step at 2 GPUs).
    if (p.world > 1 && (t->type == MS_UPDATE || p.sys_fence_all)) fence_acqrel_sys();
    for (int r = 0; r < p.world; ++r)
      if ((t->sig_mask >> r) & 1) st_relaxed_sys_u32(peers->flags[r] + t->sig_flag * 8 + p.rank, epoch);
  }
}

// warp 2 of the epilogue group: fetch this step's early-stop decision (patience > 0) into shared memory
ST_DEVINL void ms_fetch_decision(MsShared* sh, const MsParams& p, int ti, uint32_t epoch, int lane) {
  if (lane == 0) {
    uint32_t stop = 0;
    if (p.patience > 0) {
      const uint32_t* d = p.counters + p.decide_cnt * MS_CNT_STRIDE;
      uint32_t val, spins = 0;
      unsigned long long t0 = 0;
      while (int32_t(((val = ld_relaxed_gpu_u32(d)) >> 1) - epoch) < 0) {
        if ((++spins & 0x3ff) == 0) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) ms_timeout(p.error, 3, ti, epoch);
        }
      }
      fence_acqrel_gpu();
      stop = val & 1u;
    }
    sh->stop = stop;
  }
  __syncwarp();
}

// Everything the optimizer needs, by value (registers): kernel parameters must not be reached through a reference from
// a non-inlined function - that turns every field access into a generic load from parameter memory.
struct MsOptConsts {
  float bc1_inv, bc2_inv_sqrt;
  bool first;
  int kind;
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  int nesterov, maximize;
};

ST_DEVINL MsOptConsts ms_opt_consts(const MsParams& p, int64_t step1) {
  MsOptConsts oc{1.f, 1.f, step1 == 1, p.kind, p.lr, p.beta1, p.beta2, p.eps, p.weight_decay, p.momentum, p.dampening, p.nesterov, p.maximize};
  if (p.kind != 0) {
    // beta^t = 2^(t log2 beta) on the SFU (two MUFU ops instead of the ~300-instruction powf; |error| << 1e-6 here)
    oc.bc1_inv = 1.f / (1.f - exp2f(float(step1) * __log2f(p.beta1)));
    oc.bc2_inv_sqrt = rsqrtf(1.f - exp2f(float(step1) * __log2f(p.beta2)));
  }
  return oc;
}

ST_DEVINL float ms_opt_one(const MsOptConsts& a, const MsOptConsts& oc, float p, float g, float& m, float& v) {
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
  float sq;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));      // SFU square root (handles v = 0 without a slow path)
  const float denom = sq * oc.bc2_inv_sqrt + a.eps;
  return p - a.lr * oc.bc1_inv * __fdividef(m, denom);
}

ST_DEVINL void ms_opt4(const MsOptConsts& oc, float4& p, const float4& g, float4& m, float4& v) {
  p.x = ms_opt_one(oc, oc, p.x, g.x, m.x, v.x);
  p.y = ms_opt_one(oc, oc, p.y, g.y, m.y, v.y);
  p.z = ms_opt_one(oc, oc, p.z, g.z, m.z, v.z);
  p.w = ms_opt_one(oc, oc, p.w, g.w, m.w, v.w);
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

// fp32 master + bf16 operand copy of 4 consecutive updated parameters -> every replica
ST_DEVINL void ms_publish4(const MsParams& p, const MsPeerTable* peers, const ParamSegment& sg, int64_t e, const float4& pq) {
  if (p.world == 1) {
    *reinterpret_cast<float4*>(p.param_local + e) = pq;
  } else if (p.mc_param != nullptr) {
    mc_st_f4(p.mc_param + e, pq);
  } else {
    for (int r = 0; r < p.world; ++r) *reinterpret_cast<float4*>(peers->param[r] + e) = pq;
  }
  if (sg.w_off < 0) return;
  const int64_t local = e - sg.offset;
  const int64_t count = int64_t(sg.rows) * sg.cols;
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
#pragma unroll 1
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

// ---------------------------------------------------------------------------------------------- epilogues
// Rolled epilogue: 8 accumulator columns per iteration (thread = output row).  Handles bias, ReLU, ReLU-backward mask,
// bf16 / fp32 stores, per-tile column sums and - on one GPU - the optimizer fused into the weight-gradient tile.
__device__ __noinline__ void ms_epilogue_rolled(const MsTask* t, MsShared* sh, const MsParams& p, int par, uint32_t tacc, int q, int lane,
                                                uint64_t* tmem_empty, uint64_t* tmem_full, uint32_t full_parity, int64_t step1,
                                                long long* stamp_slot, int half) {
  const int flags = t->flags;
  const int row = t->m0 + q * 32 + lane;
  const bool row_ok = row < t->M;
  // the two epilogue warp groups split the tile's columns in whole 8-column groups
  const int w0 = t->ep_c0, wc = t->ep_cols ? t->ep_cols : t->block_n;     // epilogue window of this task inside the tile
  const int ncols_all = min(wc, t->N - t->n0 - w0);      // valid columns of the window
  const int c_split = ((max(ncols_all, 0) + 15) >> 4) << 3;
  const int c_rel = half ? c_split : 0;
  const int c_begin = w0 + c_rel;
  const int ncols = (half ? ncols_all : min(c_split, ncols_all)) - c_rel;       // this group's columns (may be <= 0)
  const int n0 = t->n0 + c_begin;
  tacc += uint32_t(c_begin);
  const bool have_mask = (flags & MS_MASK) != 0;
  const bool fused = (flags & MS_FUSED_OPT) != 0 && sh->stop == 0;
  const __nv_bfloat16* mrow = have_mask ? reinterpret_cast<const __nv_bfloat16*>(t->mask[par]) + size_t(row_ok ? row : 0) * t->ld_mask + n0 : nullptr;
  __nv_bfloat16* orow = (flags & MS_OUT_BF16) ? reinterpret_cast<__nv_bfloat16*>(t->out_bf16[par]) + size_t(row) * t->ld_out + n0 : nullptr;
  float* frow = (flags & MS_OUT_F32) ? reinterpret_cast<float*>(t->out_f32) + size_t(row) * t->ld_f32 + n0 : nullptr;
  const bool f32_vec = (t->ld_f32 & 3) == 0;
  // fused optimizer: the tile's output IS the gradient of parameters e_row .. (arena elements), updated in place
  const int64_t e_row = (flags & MS_FUSED_OPT) ? (reinterpret_cast<float*>(t->out_f32) - p.grad_local) + int64_t(row) * t->ld_f32 + n0 : 0;
  const bool has_m = (p.kind != 0) || (p.momentum != 0.f);
  MsOptConsts oc{};
  ParamSegment sg{};
  if (fused) {
    oc = ms_opt_consts(p, step1);
    sg = p.segs[t->seg];
  }
  uint4 mk_next = make_uint4(0, 0, 0, 0);
  if (have_mask && row_ok && ncols >= 8) mk_next = ld_cg_u4(mrow);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 pn[2] = {z4, z4}, mn[2] = {z4, z4}, vn[2] = {z4, z4};   // next iteration's parameters / optimizer state (prefetched)
  auto load_opt = [&](int c) {
    if (fused && row_ok && f32_vec && c + 8 <= ncols) {
      const int64_t e = e_row + c;
      pn[0] = ld_cg_f4(p.param_local + e); pn[1] = ld_cg_f4(p.param_local + e + 4);
      if (has_m) { mn[0] = ld_cg_f4(p.m + e - p.shard_lo); mn[1] = ld_cg_f4(p.m + e + 4 - p.shard_lo); }
      if (p.kind != 0) { vn[0] = ld_cg_f4(p.v + e - p.shard_lo); vn[1] = ld_cg_f4(p.v + e + 4 - p.shard_lo); }
    }
  };
  load_opt(0);
  // everything above (first mask chunk, first optimizer operands) is in flight while the main loop of this tile still runs
  ms_mbar_wait(tmem_full, full_parity, p, 13);
  tc_fence_after();
  if (stamp_slot != nullptr) *stamp_slot = (long long)globaltimer_ns();
  if (ncols <= 0) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty);
    return;
  }
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 8) {
    uint32_t r[8];
    tmem_ld_32x8(tacc + uint32_t(c0), r);
    const uint4 mk = mk_next;
    if (have_mask && row_ok && c0 + 16 <= ncols) mk_next = ld_cg_u4(mrow + c0 + 8);
    float4 pc[2] = {pn[0], pn[1]}, mc[2] = {mn[0], mn[1]}, vc[2] = {vn[0], vn[1]};
    load_opt(c0 + 8);
    tmem_ld_wait();
    if (c0 + 8 >= ncols) {
      // last TMEM read of this tile: hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
    const bool full = c0 + 8 <= ncols;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = __uint_as_float(r[j]);
      if (flags & MS_BIAS) v += sh->bias[c_begin + c0 + j];
      if (flags & MS_RELU) v = fmaxf(v, 0.f);
      a[j] = v;
    }
    if (have_mask && row_ok) {
      if (full) {
        const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float mv = __uint_as_float((j & 1) ? (w[j >> 1] & 0xffff0000u) : (w[j >> 1] << 16));
          a[j] = (mv > 0.f) ? a[j] : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < ncols) {
            const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(mrow + c0 + j));
            a[j] = (__uint_as_float(uint32_t(raw) << 16) > 0.f) ? a[j] : 0.f;
          }
      }
    }
    if (!row_ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = 0.f;
    }
    if (orow != nullptr && row_ok) {
      if (full) {
        uint4 w;
        w.x = pack_bf16x2(a[0], a[1]); w.y = pack_bf16x2(a[2], a[3]);
        w.z = pack_bf16x2(a[4], a[5]); w.w = pack_bf16x2(a[6], a[7]);
        *reinterpret_cast<uint4*>(orow + c0) = w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < ncols) orow[c0 + j] = __float2bfloat16(a[j]);
      }
    }
    if (frow != nullptr && row_ok) {
      if (fused && full && f32_vec) {
        // ---- optimizer on the 8 parameters whose gradient this thread holds; master, state and bf16 operand copy updated in place
        const float4 g0 = make_float4(a[0], a[1], a[2], a[3]), g1 = make_float4(a[4], a[5], a[6], a[7]);
        ms_opt4(oc, pc[0], g0, mc[0], vc[0]);
        ms_opt4(oc, pc[1], g1, mc[1], vc[1]);
        const int64_t e = e_row + c0;
        *reinterpret_cast<float4*>(p.param_local + e) = pc[0];
        *reinterpret_cast<float4*>(p.param_local + e + 4) = pc[1];
        if (has_m) { *reinterpret_cast<float4*>(p.m + e - p.shard_lo) = mc[0]; *reinterpret_cast<float4*>(p.m + e + 4 - p.shard_lo) = mc[1]; }
        if (p.kind != 0) { *reinterpret_cast<float4*>(p.v + e - p.shard_lo) = vc[0]; *reinterpret_cast<float4*>(p.v + e + 4 - p.shard_lo) = vc[1]; }
        if (sg.w_off >= 0) {
          uint4 w;
          w.x = pack_bf16x2(pc[0].x, pc[0].y); w.y = pack_bf16x2(pc[0].z, pc[0].w);
          w.z = pack_bf16x2(pc[1].x, pc[1].y); w.w = pack_bf16x2(pc[1].z, pc[1].w);
          *reinterpret_cast<uint4*>(p.wbf16_local + sg.w_off + int64_t(row) * sg.w_ld + n0 + c0) = w;
        }
      } else if (fused) {
        // ragged tail / unaligned pitch: element by element
#pragma unroll 1
        for (int j = 0; j < 8; ++j)
          if (c0 + j < ncols) {
            const int64_t e = e_row + c0 + j;
            float pe = ld_cg_f32(p.param_local + e), me = has_m ? ld_cg_f32(p.m + e - p.shard_lo) : 0.f;
            float ve = (p.kind != 0) ? ld_cg_f32(p.v + e - p.shard_lo) : 0.f;
            float gj = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) gj = (k == j) ? a[k] : gj;
            pe = ms_opt_one(oc, oc, pe, gj, me, ve);
            p.param_local[e] = pe;
            if (has_m) p.m[e - p.shard_lo] = me;
            if (p.kind != 0) p.v[e - p.shard_lo] = ve;
            if (sg.w_off >= 0) p.wbf16_local[sg.w_off + int64_t(row) * sg.w_ld + n0 + c0 + j] = __float2bfloat16(pe);
          }
      } else if (!(flags & MS_FUSED_OPT)) {
        if (full && f32_vec) {
          *reinterpret_cast<float4*>(frow + c0) = make_float4(a[0], a[1], a[2], a[3]);
          *reinterpret_cast<float4*>(frow + c0 + 4) = make_float4(a[4], a[5], a[6], a[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + j < ncols) frow[c0 + j] = a[j];
        }
      }
    }
    if (flags & MS_COLSUM) {
      // transposing butterfly: after the folds lane L holds the sum over the warp's 32 rows of column (L & 7)
      float v4[4], v2[2], v1;
      {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) v4[i] = (up ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, up ? a[i] : a[i + 4], 4);
      }
      {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) v2[i] = (up ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, up ? v4[i] : v4[i + 2], 2);
      }
      {
        const bool up = (lane & 1) != 0;
        v1 = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, up ? v2[0] : v2[1], 1);
      }
      v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
      if (lane < 8) sh->colsum[q][c_begin + c0 + lane] = v1;
    }
  }
}

// ---------------------------------------------------------------------------------------------- lean epilogues
// The hot tiles have whole 8-column groups and aligned pitches; their epilogues are specialised at compile time (no
// feature tests inside the loop: the generic version below spends ~400 instructions per 8 columns on them).  The TMEM load
// of the next 8 columns is issued before the current ones are processed.
template <bool BIAS, bool RELU, bool MASK, bool COLSUM>
ST_DEVINL void ms_epi_bf16_lean(const MsTask* t, MsShared* sh, int par, uint32_t tacc, int row, bool row_ok, int lane,
                                int c_begin, int ncols, uint64_t* tmem_empty, int q) {
  const int n0 = t->n0 + c_begin;
  __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(t->out_bf16[par]) + size_t(row) * t->ld_out + n0;
  const __nv_bfloat16* mrow = nullptr;
  uint4 mk_next = make_uint4(0, 0, 0, 0);
  if constexpr (MASK) {
    mrow = reinterpret_cast<const __nv_bfloat16*>(t->mask[par]) + size_t(row_ok ? row : 0) * t->ld_mask + n0;
    if (row_ok) mk_next = ld_cg_u4(mrow);
  }
  uint32_t rn[8];
  tmem_ld_32x8(tacc + uint32_t(c_begin), rn);
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 8) {
    tmem_ld_wait();
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = rn[j];
    const uint4 mk = mk_next;
    if (c0 + 8 < ncols) {
      tmem_ld_32x8(tacc + uint32_t(c_begin + c0 + 8), rn);
      if constexpr (MASK) { if (row_ok) mk_next = ld_cg_u4(mrow + c0 + 8); }
    } else {
      tc_fence_before();                         // last TMEM read of this warp: hand the accumulator stage back
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
    float a[8];
    if constexpr (BIAS) {
      const float4 b0 = *reinterpret_cast<const float4*>(&sh->bias[c_begin + c0]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sh->bias[c_begin + c0 + 4]);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __uint_as_float(r[j]) + bb[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __uint_as_float(r[j]);
    }
    if constexpr (RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
    }
    if constexpr (MASK) {
      const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float mv = __uint_as_float((j & 1) ? (w[j >> 1] & 0xffff0000u) : (w[j >> 1] << 16));
        a[j] = (mv > 0.f) ? a[j] : 0.f;
      }
    }
    if (row_ok) {
      uint4 w;
      w.x = pack_bf16x2(a[0], a[1]); w.y = pack_bf16x2(a[2], a[3]);
      w.z = pack_bf16x2(a[4], a[5]); w.w = pack_bf16x2(a[6], a[7]);
      *reinterpret_cast<uint4*>(orow + c0) = w;
    }
    if constexpr (COLSUM) {
      if (!row_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
      }
      float v4[4], v2[2], v1;
      {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) v4[i] = (up ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, up ? a[i] : a[i + 4], 4);
      }
      {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) v2[i] = (up ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, up ? v4[i] : v4[i + 2], 2);
      }
      {
        const bool up = (lane & 1) != 0;
        v1 = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, up ? v2[0] : v2[1], 1);
      }
      v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
      if (lane < 8) sh->colsum[q][c_begin + c0 + lane] = v1;
    }
  }
}

// weight-gradient tile: plain fp32 stores into the gradient arena, or (one GPU) the optimizer applied in place
template <bool FUSED>
ST_DEVINL void ms_epi_f32_lean(const MsTask* t, const MsParams& p, uint32_t tacc, int row, bool row_ok, int lane, int c_begin, int ncols,
                               uint64_t* tmem_empty, int64_t step1) {
  const int n0 = t->n0 + c_begin;
  float* frow = reinterpret_cast<float*>(t->out_f32) + size_t(row) * t->ld_f32 + n0;
  const int64_t e_row = (reinterpret_cast<float*>(t->out_f32) - p.grad_local) + int64_t(row) * t->ld_f32 + n0;
  const bool has_m = (p.kind != 0) || (p.momentum != 0.f);
  const bool has_v = p.kind != 0;
  float* P = p.param_local + e_row;
  float* Mq = p.m + (e_row - p.shard_lo);
  float* Vq = p.v + (e_row - p.shard_lo);
  __nv_bfloat16* wrow = nullptr;
  MsOptConsts oc{};
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 pn[2] = {z4, z4}, mn[2] = {z4, z4}, vn[2] = {z4, z4};
  if constexpr (FUSED) {
    oc = ms_opt_consts(p, step1);
    const ParamSegment sg = p.segs[t->seg];
    if (sg.w_off >= 0) wrow = p.wbf16_local + sg.w_off + int64_t(row) * sg.w_ld + n0;
    if (row_ok) {
      pn[0] = ld_cg_f4(P); pn[1] = ld_cg_f4(P + 4);
      if (has_m) { mn[0] = ld_cg_f4(Mq); mn[1] = ld_cg_f4(Mq + 4); }
      if (has_v) { vn[0] = ld_cg_f4(Vq); vn[1] = ld_cg_f4(Vq + 4); }
    }
  }
  uint32_t rn[8];
  tmem_ld_32x8(tacc + uint32_t(c_begin), rn);
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 8) {
    tmem_ld_wait();
    float4 g0 = make_float4(__uint_as_float(rn[0]), __uint_as_float(rn[1]), __uint_as_float(rn[2]), __uint_as_float(rn[3]));
    float4 g1 = make_float4(__uint_as_float(rn[4]), __uint_as_float(rn[5]), __uint_as_float(rn[6]), __uint_as_float(rn[7]));
    float4 pc[2] = {pn[0], pn[1]}, mc[2] = {mn[0], mn[1]}, vc[2] = {vn[0], vn[1]};
    if (c0 + 8 < ncols) {
      tmem_ld_32x8(tacc + uint32_t(c_begin + c0 + 8), rn);
      if constexpr (FUSED) {
        if (row_ok) {
          pn[0] = ld_cg_f4(P + c0 + 8); pn[1] = ld_cg_f4(P + c0 + 12);
          if (has_m) { mn[0] = ld_cg_f4(Mq + c0 + 8); mn[1] = ld_cg_f4(Mq + c0 + 12); }
          if (has_v) { vn[0] = ld_cg_f4(Vq + c0 + 8); vn[1] = ld_cg_f4(Vq + c0 + 12); }
        }
      }
    } else {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
    if (!row_ok) continue;
    if constexpr (FUSED) {
      ms_opt4(oc, pc[0], g0, mc[0], vc[0]);
      ms_opt4(oc, pc[1], g1, mc[1], vc[1]);
      *reinterpret_cast<float4*>(P + c0) = pc[0];
      *reinterpret_cast<float4*>(P + c0 + 4) = pc[1];
      if (has_m) { *reinterpret_cast<float4*>(Mq + c0) = mc[0]; *reinterpret_cast<float4*>(Mq + c0 + 4) = mc[1]; }
      if (has_v) { *reinterpret_cast<float4*>(Vq + c0) = vc[0]; *reinterpret_cast<float4*>(Vq + c0 + 4) = vc[1]; }
      if (wrow != nullptr) {
        uint4 w;
        w.x = pack_bf16x2(pc[0].x, pc[0].y); w.y = pack_bf16x2(pc[0].z, pc[0].w);
        w.z = pack_bf16x2(pc[1].x, pc[1].y); w.w = pack_bf16x2(pc[1].z, pc[1].w);
        *reinterpret_cast<uint4*>(wrow + c0) = w;
      }
    } else {
      *reinterpret_cast<float4*>(frow + c0) = g0;
      *reinterpret_cast<float4*>(frow + c0 + 4) = g1;
    }
  }
}

// picks the specialised epilogue for a tile; false: the tile is ragged / unaligned -> generic version
ST_DEVINL bool ms_epilogue_lean(const MsTask* t, MsShared* sh, const MsParams& p, int par, uint32_t tacc, int q, int lane,
                                uint64_t* tmem_empty, uint64_t* tmem_full, uint32_t full_parity, int64_t step1, long long* stamp_slot,
                                int half) {
  const int flags = t->flags;
  const int w0 = t->ep_c0, wc = t->ep_cols ? t->ep_cols : t->block_n;     // epilogue window of this task inside the tile
  const int ncols_all = min(wc, t->N - t->n0 - w0);
  if (ncols_all <= 0 || (ncols_all & 7) != 0 || (w0 & 7) != 0) return false;
  if ((flags & MS_OUT_F32) && (((t->ld_f32 & 3) != 0) || (flags & MS_OUT_BF16))) return false;
  const int c_split = ((ncols_all + 15) >> 4) << 3;
  const int c_rel = half ? c_split : 0;
  const int c_begin = w0 + c_rel;
  const int ncols = (half ? ncols_all : min(c_split, ncols_all)) - c_rel;
  const int row = t->m0 + q * 32 + lane;
  const bool row_ok = row < t->M;
  ms_mbar_wait(tmem_full, full_parity, p, 13);
  tc_fence_after();
  if (stamp_slot != nullptr) *stamp_slot = (long long)globaltimer_ns();
  if (ncols <= 0) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty);
    return true;
  }
  if (flags & MS_OUT_F32) {
    if ((flags & MS_FUSED_OPT) && sh->stop == 0) ms_epi_f32_lean<true>(t, p, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, step1);
    else if (flags & MS_FUSED_OPT) {               // stopped: the update is skipped, nothing to store
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    } else ms_epi_f32_lean<false>(t, p, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, step1);
    return true;
  }
  const int sel = ((flags & MS_BIAS) ? 1 : 0) | ((flags & MS_RELU) ? 2 : 0) | ((flags & MS_MASK) ? 4 : 0) | ((flags & MS_COLSUM) ? 8 : 0);
  switch (sel) {
    case 0: ms_epi_bf16_lean<false, false, false, false>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    case 1: ms_epi_bf16_lean<true, false, false, false>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    case 2: ms_epi_bf16_lean<false, true, false, false>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    case 3: ms_epi_bf16_lean<true, true, false, false>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    case 4: ms_epi_bf16_lean<false, false, true, false>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    case 8: ms_epi_bf16_lean<false, false, false, true>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    case 12: ms_epi_bf16_lean<false, false, true, true>(t, sh, par, tacc, row, row_ok, lane, c_begin, ncols, tmem_empty, q); break;
    default: return false;                          // (never reached for the tasks mega_mlp.py builds; keeps the contract honest)
  }
  return true;
}

// Head tile with the fused loss: the thread owns the whole output row (N <= NC columns, one TMEM load).  NC = 16 for the usual
// 10-class heads: half the exponentials / shuffles of the 32-wide version.
template <int NC>
ST_DEVINL void ms_epilogue_loss(const MsTask* t, MsShared* sh, const MsParams& p, int par, uint32_t tacc, int q, int lane,
                                int lbl_off, uint64_t* tmem_empty, uint64_t* tmem_full, uint32_t full_parity,
                                long long* stamp_slot, int half, uint8_t* chain) {
  if (half) {                                    // the row-wise loss is done by the first warp group alone
    ms_mbar_wait(tmem_full, full_parity, p, 13);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty);
    return;
  }
  const int flags = t->flags;
  const int N = t->N;
  const int row = t->m0 + q * 32 + lane;
  const bool row_ok = row < t->M;
  int y = 0;
  if ((flags & MS_LOSS_CE) && row_ok) y = reinterpret_cast<const int32_t*>(t->labels)[row + lbl_off];
  ms_mbar_wait(tmem_full, full_parity, p, 13);
  tc_fence_after();
  if (stamp_slot != nullptr) *stamp_slot = (long long)globaltimer_ns();
  uint32_t r[32];
  if constexpr (NC == 32) tmem_ld_32x32(tacc, r); else tmem_ld_32x16(tacc, r);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty);
  float acc[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float v = __uint_as_float(r[j]);
    if (flags & MS_BIAS) v += sh->bias[j];
    if (flags & MS_RELU) v = fmaxf(v, 0.f);
    acc[j] = v;
  }
  float li = 0.f;
  if (flags & MS_LOSS_CE) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NC; ++j) if (j < N) mx = fmaxf(mx, acc[j]);
    float sum = 0.f, zy = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (j == y) zy = acc[j];
      acc[j] = (j < N) ? __expf(acc[j] - mx) : 0.f;
      sum += acc[j];
    }
    li = row_ok ? (mx + __logf(sum) - zy) : 0.f;
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = (row_ok && j < N) ? (acc[j] * inv - (j == y ? 1.f : 0.f)) * t->loss_scale : 0.f;
  } else {
    const float* target = reinterpret_cast<const float*>(t->labels);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      float dlt = 0.f;
      if (row_ok && j < N) dlt = acc[j] - target[size_t(row + lbl_off) * t->ld_target + j];
      li += dlt * dlt;
      acc[j] = 2.f * dlt * t->loss_scale;
    }
  }
  if (!(flags & MS_QUIET)) {
#pragma unroll
    for (int o = 16; o; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
    if (lane == 0) sh->loss_w[q] = li * t->loss_scale;
  }
  // dZ_L (bf16): whole 8-column groups up to the pitch (a multiple of 8); columns >= N inside them are written as zeros
  {
    const bool to_global = row_ok && !(flags & MS_QUIET);
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(t->out_bf16[par]) + size_t(row) * t->ld_out;
    const int rl = q * 32 + lane;              // row inside the tile = row of the chain buffer
#pragma unroll
    for (int v = 0; v < NC / 8; ++v) {
      uint4 w;
      w.x = pack_bf16x2(acc[8 * v + 0], acc[8 * v + 1]); w.y = pack_bf16x2(acc[8 * v + 2], acc[8 * v + 3]);
      w.z = pack_bf16x2(acc[8 * v + 4], acc[8 * v + 5]); w.w = pack_bf16x2(acc[8 * v + 6], acc[8 * v + 7]);
      if (to_global && 8 * v < N && 8 * v + 8 <= t->ld_out) reinterpret_cast<uint4*>(op)[v] = w;
      // chain buffer: K-major operand with the 128-byte swizzle the MMA descriptor expects (16-byte chunk c of row r sits
      // at chunk c ^ (r & 7)); rows past M and columns past N carry zeros (acc is zero there)
      if (chain != nullptr) *reinterpret_cast<uint4*>(chain + rl * 128 + ((v ^ (rl & 7)) << 4)) = w;
    }
    if (chain != nullptr) {
      fence_proxy_async_all();                 // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->chain_bar);
    }
  }
  if ((flags & MS_COLSUM) && !(flags & MS_QUIET)) {
    // butterfly over the 32 rows of the warp: afterwards lane L holds the sum of column (L mod NC)
    float v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = acc[j];
#pragma unroll
    for (int s = NC / 2; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float send = up ? v[i] : v[i + s];
        const float keep = up ? v[i + s] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    if constexpr (NC == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
    if (lane < NC) sh->colsum[q][lane] = v[0];
  }
}

// Owner-side reduce + optimizer + publication of one chunk of the flat arena (256 threads, two float4 per thread in flight).
ST_DEVINL void ms_update_chunk(const MsTask* t, const MsParams& p, const MsPeerTable* peers, int et, int64_t step1) {
  const ParamSegment sg = p.segs[t->seg];
  const MsOptConsts oc = ms_opt_consts(p, step1);
  const float inv_w = 1.f / float(p.world);
  const bool has_m = (p.kind != 0) || (p.momentum != 0.f);
  const int n4 = t->len >> 2;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int base = 0; base < n4; base += 2 * MS_EPI_THREADS) {
    float4 g[2], pp[2], mm[2], vv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {              // all loads of both elements first: one NVLink / L2 round trip
      const int i4 = base + u * MS_EPI_THREADS + et;
      g[u] = pp[u] = mm[u] = vv[u] = z4;
      if (i4 < n4) {
        const int64_t e = t->e0 + (int64_t(i4) << 2);
        g[u] = ms_reduce4(p, peers, e);
        for (int pr = 1; pr < t->n_part; ++pr) {
          const float4 x = ms_reduce4(p, peers, t->part_off + int64_t(pr - 1) * t->part_stride + e);
          g[u].x += x.x; g[u].y += x.y; g[u].z += x.z; g[u].w += x.w;
        }
        pp[u] = ld_cg_f4(p.param_local + e);
        if (has_m) mm[u] = ld_cg_f4(p.m + (e - p.shard_lo));
        if (p.kind != 0) vv[u] = ld_cg_f4(p.v + (e - p.shard_lo));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i4 = base + u * MS_EPI_THREADS + et;
      if (i4 >= n4) continue;
      const int64_t e = t->e0 + (int64_t(i4) << 2);
      float4 gq = g[u];
      gq.x *= inv_w; gq.y *= inv_w; gq.z *= inv_w; gq.w *= inv_w;
      ms_opt4(oc, pp[u], gq, mm[u], vv[u]);
      if (has_m) *reinterpret_cast<float4*>(p.m + (e - p.shard_lo)) = mm[u];
      if (p.kind != 0) *reinterpret_cast<float4*>(p.v + (e - p.shard_lo)) = vv[u];
      ms_publish4(p, peers, sg, e, pp[u]);
    }
  }
}

// `loss` = the replicas' loss partials summed in (rank, m-tile) order, gathered by the whole warp (see the call site)
ST_DEVINL void ms_decide(const MsTask* t, const MsParams& p, const MsPeerTable* peers, uint32_t epoch, int64_t it, float loss) {
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
  if (p.patience > 0) st_release_gpu_u32(p.counters + p.decide_cnt * MS_CNT_STRIDE, epoch * 2u + stop_bit);
  ms_signal(t, peers, p, epoch);
}

}  // namespace

__global__ void __launch_bounds__(MS_THREADS, 1) mlp_step_kernel(const __grid_constant__ MsParams p) {
  // iter0 < 0: take the step numbers from the device-side state (what the previous launch left there) - this is what
  // lets a CUDA graph with this kernel in it be replayed step after step (input streaming, gather + step graphs).
  int64_t iter0 = p.iter0, step0 = p.step0;
  uint32_t epoch0 = p.epoch0;
  if (p.iter0 < 0) {
    iter0 = p.state->iter;
    step0 = p.state->step;
    epoch0 = p.state->pad[0];
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* chain = smem + MS_CHAIN_OFF;
  MsShared* sh = reinterpret_cast<MsShared*>(smem + MS_MISC_OFF);
  MsPeerTable* peers = reinterpret_cast<MsPeerTable*>(smem + MS_MISC_OFF + MS_MISC_BYTES);
  MsTask* s_tasks = reinterpret_cast<MsTask*>(smem + MS_MISC_OFF + MS_MISC_BYTES + sizeof(MsPeerTable));

  const int warp = ubc(threadIdx.x >> 5);
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
      mbar_init(&sh->tmem_empty_bar[a], MS_EPI_THREADS / 32);   // one arrive per epilogue warp
    }
    mbar_init(&sh->chain_bar, 4);
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
    if (threadIdx.x < 64) {
      sh->bias[threadIdx.x] = 0.f;
      for (int k = 0; k < 4; ++k) sh->colsum[k][threadIdx.x] = 0.f;
    }
    if (threadIdx.x < 4) sh->loss_w[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) sh->stop = 0;
    // chain buffer: only the 16-byte chunks that hold dZ_L are ever written again, the rest stays zero (K padding of the MMA)
    for (int i = threadIdx.x; i < MS_CHAIN_BYTES / 16; i += MS_THREADS) reinterpret_cast<uint4*>(chain)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_all();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uint32_t(ubc(int(sh->tmem_ptr)));
  const uint32_t smem_base = smem_u32(smem);
  long long* prof = p.prof ? p.prof + size_t(blockIdx.x) * MS_MAX_TASKS * 8 : nullptr;
  long long* prof_kb = p.prof_kb ? p.prof_kb + size_t(blockIdx.x) * MS_MAX_TASKS * 64 : nullptr;

  if (warp == 0) {
    // ===================== dependency waits + TMA producer (whole warp runs the loops, one elected lane issues) ==========
    uint32_t kbc = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const uint32_t epoch = epoch0 + uint32_t(s) + 1u;
      const int par = int((epoch0 + uint32_t(s)) & 1u);
      const int win_off = int(((iter0 + s) % p.n_windows) * p.window_rows);
      const bool stamp = prof != nullptr && s == p.n_steps - 1;
      for (int ti = 0; ti < n_tasks; ++ti) {
        const MsTask* t = &s_tasks[ti];
        if (ubc(t->type) != MS_GEMM) continue;
        const int flags = ubc(t->flags);
        const CUtensorMap* ma = p.tmaps + ubc(t->tmap_a[par]);
        const CUtensorMap* mb = p.tmaps + ubc(t->tmap_b[par]);
        const bool a_mn = (flags & MS_A_MN) != 0, b_mn = (flags & MS_B_MN) != 0;
        const int a_row_off = (flags & MS_A_WIN) ? win_off : 0;
        const int b_k_off = (flags & MS_B_WIN) ? win_off : 0;
        const int m0 = ubc(t->m0), n0 = ubc(t->n0), nkb = ubc(t->num_kb);
        const bool chain_tail = (flags & MS_CHAIN_TAIL) != 0;     // A comes from the chain buffer: only B travels
        const uint32_t tx = (chain_tail ? 0u : uint32_t(MS_A_BYTES)) + (b_mn ? uint32_t(MS_B_BYTES) : uint32_t(ubc(t->block_n)) * 128u);
        const int pre = nkb < MS_STAGES ? nkb : MS_STAGES;      // stages armed before the late dependencies are met
        if (elect_one()) {
          if (stamp) prof[ti * 8 + 0] = (long long)globaltimer_ns();
          tma_prefetch_desc(ma);                   // descriptor fetch overlaps the dependency wait
          tma_prefetch_desc(mb);
        }
        __syncwarp();
        // -- early: what the early operand needs (B: weights / the resident batch; MS_EARLY_A tiles: nothing - their A is the input
        //    batch), then that operand for the first `pre` k-blocks
        ms_wait_warp(t, ti, p, epoch, lane, 1);
        fence_proxy_async_all();
        if (stamp && elect_one()) prof[ti * 8 + 1] = (long long)globaltimer_ns();
        const bool early_a = (flags & MS_EARLY_A) != 0;           // (K-major A only: the batch window of the first layer)
        for (int i = 0; i < pre; ++i) {
          const uint32_t kk = kbc + uint32_t(i);
          const uint32_t st_ = kk & (MS_STAGES - 1);
          ms_mbar_wait(&sh->empty_bar[st_], ((kk / MS_STAGES) & 1) ^ 1, p, 10);
          if (elect_one()) {
            uint8_t* sa = smem + st_ * MS_STAGE_BYTES;
            uint8_t* sb = sa + MS_A_BYTES;
            mbar_arrive_expect_tx(&sh->full_bar[st_], tx);           // both operands' bytes: the stage completes with the second one
            const int k0 = i * MS_BLOCK_K;
            if (early_a) tma_load_2d(sa, ma, &sh->full_bar[st_], k0, m0 + a_row_off);
            else if (b_mn) tma_load_2d(sb, mb, &sh->full_bar[st_], n0, k0 + b_k_off);
            else tma_load_2d(sb, mb, &sh->full_bar[st_], k0, n0);
            if (stamp && prof_kb && i < 16) prof_kb[ti * 64 + i * 4 + 0] = (long long)globaltimer_ns();
          }
          __syncwarp();
        }
        // -- late: the A operand (what the previous phase of this step produced)
        ms_wait_warp(t, ti, p, epoch, lane, 0);
        fence_proxy_async_all();                   // other CTAs' generic-proxy stores -> visible to the TMA (async proxy) loads below
        if (stamp && elect_one()) prof[ti * 8 + 2] = (long long)globaltimer_ns();
        for (int i = 0; i < nkb; ++i, ++kbc) {
          const uint32_t st_ = kbc & (MS_STAGES - 1);
          if (i >= pre) ms_mbar_wait(&sh->empty_bar[st_], ((kbc / MS_STAGES) & 1) ^ 1, p, 10);
          if (elect_one()) {
            uint8_t* sa = smem + st_ * MS_STAGE_BYTES;
            uint8_t* sb = sa + MS_A_BYTES;
            const int k0 = i * MS_BLOCK_K;
            if (i >= pre) mbar_arrive_expect_tx(&sh->full_bar[st_], tx);
            if (i >= pre || early_a) {
              if (b_mn) tma_load_2d(sb, mb, &sh->full_bar[st_], n0, k0 + b_k_off);
              else tma_load_2d(sb, mb, &sh->full_bar[st_], k0, n0);
            }
            if (chain_tail || (early_a && i < pre)) {
            } else if (a_mn) {
              tma_load_2d(sa, ma, &sh->full_bar[st_], m0, k0);
              tma_load_2d(sa + MS_BLOCK_K * 128, ma, &sh->full_bar[st_], m0 + 64, k0);
            } else {
              tma_load_2d(sa, ma, &sh->full_bar[st_], k0, m0 + a_row_off);
            }
            if (stamp && prof_kb && i < 16) prof_kb[ti * 64 + i * 4 + 1] = (long long)globaltimer_ns();
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp waits, one elected lane issues) =====================
    uint32_t kbc = 0, gc = 0, cc = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const bool stamp = prof != nullptr && s == p.n_steps - 1;
      for (int ti = 0; ti < n_tasks; ++ti) {
        const MsTask* t = &s_tasks[ti];
        if (ubc(t->type) != MS_GEMM) continue;
        const int flags = ubc(t->flags);
        const uint32_t as = gc & 1;
        ms_mbar_wait(&sh->tmem_empty_bar[as], ((gc >> 1) & 1) ^ 1, p, 11);
        const bool chain_tail = (flags & MS_CHAIN_TAIL) != 0;
        if (chain_tail) {                       // the head's epilogue warps have put dZ_L into the chain buffer
          ms_mbar_wait(&sh->chain_bar, cc & 1, p, 14);
          ++cc;
        }
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * MS_ACC_COLS;
        const bool a_mn = (flags & MS_A_MN) != 0, b_mn = (flags & MS_B_MN) != 0;
        const int bn_raw = ubc(t->block_n);
        const uint32_t bn = uint32_t(bn_raw < 16 ? 16 : bn_raw);
        const uint32_t idesc = make_idesc_bf16(MS_BLOCK_M, bn) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
        const uint64_t ka = a_mn ? 128 : 2, kb = b_mn ? 128 : 2;
        const int nkb = ubc(t->num_kb);
        for (int i = 0; i < nkb; ++i, ++kbc) {
          const uint32_t st_ = kbc & (MS_STAGES - 1);
          ms_mbar_wait(&sh->full_bar[st_], (kbc / MS_STAGES) & 1, p, 12);
          tc_fence_after();
          const uint32_t sa = smem_base + st_ * MS_STAGE_BYTES;
          const uint32_t sb = sa + MS_A_BYTES;
          const uint64_t da = chain_tail ? make_kmajor_sw128_desc(smem_base + MS_CHAIN_OFF)
                                         : (a_mn ? make_mnmajor_sw128_desc(sa, MS_BLOCK_K * 128) : make_kmajor_sw128_desc(sa));
          const uint64_t db = b_mn ? make_mnmajor_sw128_desc(sb, MS_BLOCK_K * 128) : make_kmajor_sw128_desc(sb);
          if (elect_one()) {
            if (stamp) {
              if (i == 0) prof[ti * 8 + 3] = (long long)globaltimer_ns();
              if (prof_kb != nullptr && i < 16) prof_kb[ti * 64 + i * 4 + 2] = (long long)globaltimer_ns();
            }
#pragma unroll
            for (int k = 0; k < MS_BLOCK_K / 16; ++k) umma_f16(tacc, da + ka * k, db + kb * k, idesc, (i | k) != 0);
            umma_commit(&sh->empty_bar[st_]);
            if (i == nkb - 1) {
              umma_commit(&sh->tmem_full_bar[as]);
              if (stamp) prof[ti * 8 + 4] = (long long)globaltimer_ns();
            }
          }
          __syncwarp();
        }
        ++gc;
      }
    }
  } else {
    // ===================== epilogue / update / decision (warps 2..5) =====================
    const int q = warp & 3;
    const int et = threadIdx.x - 64;             // 0..MS_EPI_THREADS-1
    const int half = (warp - 2) >> 2;            // warp group: which half of a tile's columns
    uint32_t gc = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const uint32_t epoch = epoch0 + uint32_t(s) + 1u;
      const int par = int((epoch0 + uint32_t(s)) & 1u);
      const int64_t it = iter0 + s;
      const int64_t step1 = step0 + s + 1;
      const int win_off = int((it % p.n_windows) * p.window_rows);
      const bool stamp = prof != nullptr && s == p.n_steps - 1;
      for (int ti = 0; ti < n_tasks; ++ti) {
        const MsTask* t = &s_tasks[ti];
        const int type = t->type;
        if (type == MS_GEMM) {
          const uint32_t as = gc & 1;
          const int flags = t->flags;
          const int n0 = t->n0, N = t->N, bn = t->block_n;
          // The epilogue warps acquire the task's dependencies themselves (same polls as the producer warp, concurrently):
          // the bias (published by last step's update), the ReLU mask (this step's forward) and the early-stop decision
          // may then be fetched while the main loop runs, instead of after the accumulator is complete.
          // (a chained tail needs neither: its mask is the activation its head copy - the task before it on this CTA - waited for)
          if ((flags & (MS_BIAS | MS_MASK | MS_FUSED_OPT)) && !(flags & MS_CHAIN_TAIL)) {
            if (warp == 2) {
              ms_wait_warp(t, ti, p, epoch, lane, -1);
              if (flags & MS_FUSED_OPT) ms_fetch_decision(sh, p, ti, epoch, lane);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(MS_EPI_THREADS) : "memory");
          }
          if (flags & MS_BIAS) {
            if (et < bn) sh->bias[et] = (n0 + et < N) ? ld_cg_f32(reinterpret_cast<const float*>(t->bias) + n0 + et) : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(MS_EPI_THREADS) : "memory");
          }
          const uint32_t tacc = tmem_base + (uint32_t(q * 32) << 16) + as * MS_ACC_COLS;
          long long* slot5 = (stamp && et == 0) ? &prof[ti * 8 + 5] : nullptr;
          if (flags & (MS_LOSS_CE | MS_LOSS_MSE)) {
            uint8_t* cb = (flags & MS_CHAIN_HEAD) ? chain : nullptr;
            if (N <= 16) ms_epilogue_loss<16>(t, sh, p, par, tacc, q, lane, (flags & MS_LBL_WIN) ? win_off : 0, &sh->tmem_empty_bar[as],
                                              &sh->tmem_full_bar[as], (gc >> 1) & 1, slot5, half, cb);
            else ms_epilogue_loss<32>(t, sh, p, par, tacc, q, lane, (flags & MS_LBL_WIN) ? win_off : 0, &sh->tmem_empty_bar[as],
                                      &sh->tmem_full_bar[as], (gc >> 1) & 1, slot5, half, cb);
          } else if (!ms_epilogue_lean(t, sh, p, par, tacc, q, lane, &sh->tmem_empty_bar[as], &sh->tmem_full_bar[as], (gc >> 1) & 1, step1,
                                       slot5, half)) {
            ms_epilogue_rolled(t, sh, p, par, tacc, q, lane, &sh->tmem_empty_bar[as], &sh->tmem_full_bar[as], (gc >> 1) & 1, step1, slot5, half);
          }
          if ((flags & (MS_COLSUM | MS_LOSS_CE | MS_LOSS_MSE)) && !(flags & MS_QUIET)) {
            asm volatile("bar.sync 1, %0;" ::"n"(MS_EPI_THREADS) : "memory");
            const int wlo = t->ep_c0, whi = t->ep_cols ? t->ep_c0 + t->ep_cols : bn;
            if ((flags & MS_COLSUM) && et >= wlo && et < whi) {
              if (n0 + et < N)
                reinterpret_cast<float*>(t->colsum)[n0 + et] = ((sh->colsum[0][et] + sh->colsum[1][et]) + sh->colsum[2][et]) + sh->colsum[3][et];
            }
            if ((flags & (MS_LOSS_CE | MS_LOSS_MSE)) && et == 0)
              *reinterpret_cast<float*>(t->loss_slot[par]) = ((sh->loss_w[0] + sh->loss_w[1]) + sh->loss_w[2]) + sh->loss_w[3];
          }
          if (stamp && et == 0) prof[ti * 8 + 6] = (long long)globaltimer_ns();
          asm volatile("bar.sync 1, %0;" ::"n"(MS_EPI_THREADS) : "memory");      // every store of the tile has been issued by its thread
          if (et == 0) {
            ms_signal(t, peers, p, epoch);
            if (stamp) prof[ti * 8 + 7] = (long long)globaltimer_ns();
          }
          ++gc;
        } else if (type == MS_UPDATE) {
          if (stamp && et == 0) prof[ti * 8 + 0] = (long long)globaltimer_ns();
          if (warp == 2) {
            ms_wait_warp(t, ti, p, epoch, lane, -1);
            ms_fetch_decision(sh, p, ti, epoch, lane);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(MS_EPI_THREADS) : "memory");
          if (stamp && et == 0) prof[ti * 8 + 2] = (long long)globaltimer_ns();
          if (sh->stop == 0) ms_update_chunk(t, p, peers, et, step1);
          if (stamp && et == 0) prof[ti * 8 + 6] = (long long)globaltimer_ns();
          asm volatile("bar.sync 1, %0;" ::"n"(MS_EPI_THREADS) : "memory");
          if (et == 0) {
            ms_signal(t, peers, p, epoch);
            if (stamp) prof[ti * 8 + 7] = (long long)globaltimer_ns();
          }
        } else if (type == MS_DECIDE) {
          if (warp == 2) {
            if (stamp && lane == 0) prof[ti * 8 + 0] = (long long)globaltimer_ns();
            ms_wait_warp(t, ti, p, epoch, lane, -1);
            // all W x n_mt loss partials are fetched at once (one NVLink round trip instead of W x n_mt in a row: the decision
            // sits in front of the next step's first forward tile), then summed in a fixed order so every rank gets the same bits
            float loss = 0.f;
            const int n_val = p.world * p.n_mt;
            for (int base = 0; base < n_val; base += 32) {
              const int i = base + lane;
              float v = 0.f;
              if (i < n_val) v = ld_relaxed_sys_f32(peers->grad[i / p.n_mt] + p.loss_off + int64_t(par) * p.n_mt + (i % p.n_mt));
              const int cnt = min(32, n_val - base);
              for (int j = 0; j < cnt; ++j) loss += __shfl_sync(0xffffffffu, v, j);
            }
            if (lane == 0) {
              ms_decide(t, p, peers, epoch, it, loss);
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
