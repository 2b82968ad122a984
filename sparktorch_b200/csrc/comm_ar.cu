// K-AR: gradient all-reduce FUSED with scale, optimizer update, precision cast,
// parameter publication and early stopping – one kernel per training step, no
// NCCL call, no host round trip.
//
// Replaces, per iteration of the reference loop (SURVEY.md §2.7 "K-AR"):
//   * one gloo all_reduce PER PARAMETER + `grad /= world` (distributed.py:179-181)
//   * two extra scalar all_reduces + two .item() syncs for early stopping (:188-196)
//   * optimizer.step() (torch.optim foreach kernels, :198)
//   * (ours) the bf16 / transposed-bf16 weight copies the tcgen05 GEMMs consume.
//
// Algorithm (two-shot over NVLink peer memory, every GPU owns 1/W of the arena):
//   barrier A  : per-block flag exchange (st.release.sys / ld.acquire.sys on
//                peer-mapped signal pads) – "all gradients are written".
//   reduce     : owner loads its shard of the fp32 gradient arena from ALL peers
//                with 128-bit P2P loads (or one multimem.ld_reduce through the
//                NVSwitch when a multicast mapping exists) and averages.
//   update     : SGD(momentum/nesterov) / Adam / AdamW on fp32 master weights;
//                optimizer state exists only for the owned shard (ZeRO-1 layout).
//   publish    : new values are stored straight into every peer's arenas over
//                NVLink: fp32 (autograd consumers), bf16 row-major and bf16
//                transposed (K-major operands of the fwd / dgrad GEMMs).
//   barrier B  : "all parameters are published"; then the local gradient arena is
//                cleared for the next step and the last block advances StepState
//                (epoch, step, iteration, early-stopping record, loss history,
//                pinned host mailbox).
// The loss and the stop vote ride in the tail slots of the gradient arena; every
// rank reads all W copies, so all ranks take the same early-stop decision without
// another exchange.  Stop semantics follow the reference: the update of the
// stopping iteration is skipped (distributed.py:185-198).
#include "common.cuh"
#include "kernels.h"

namespace st {

constexpr int AR_THREADS = 256;
constexpr int MAX_SEGS_SMEM = 512;

ST_DEVINL int flag_index(int phase, int src_rank, int block) { return (phase * 8 + src_rank) * AR_MAX_BLOCKS + block; }

// Block-level barrier across GPUs.  Caller guarantees all threads of the block
// reach it.  Data written before it (by any thread of the block, to any GPU) is
// visible to peer blocks that pass the same barrier.
template <bool RELEASE>
ST_DEVINL void block_barrier_all_gpus(uint32_t* const* flags, int world, int rank, int phase, uint32_t epoch) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < world && t != rank) {
    uint32_t* theirs = flags[t] + flag_index(phase, rank, blockIdx.x);
    if (RELEASE) {
      // everything this block stored (to any GPU) happens-before the flag: bar.sync made it visible to
      // this thread, st.release.sys orders it before the flag for system-scope observers (one MEMBAR.SYS)
      st_release_sys_u32(theirs, epoch);
    } else {
      // nothing written by THIS kernel needs to be published (the gradients were completed by earlier
      // kernels of the stream and are already performed at device scope): a plain system-scope store
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    }
    // poll with relaxed loads (no fence per iteration), one acquire fence once the flag has arrived
    const uint32_t* mine = flags[rank] + flag_index(phase, t, blockIdx.x);
    uint32_t seen;
    do {
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    } while (int32_t(seen - epoch) < 0);
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncthreads();
}

ST_DEVINL float4 ld_reduce_mc_f4(const float* mc_ptr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_ptr)
               : "memory");
  return v;
}

ST_DEVINL void st_mc_f4(float* mc_ptr, const float4& g) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
               : "memory");
}

struct OptConsts {
  float bc1_inv, bc2_inv_sqrt;  // Adam bias corrections for this step
};

ST_DEVINL float opt_update_one(const UpdateArgs& a, const OptConsts& oc, float p, float g, float& m, float& v, bool first_step) {
  if (a.maximize) g = -g;
  if (a.kind == 0) {  // SGD
    if (a.weight_decay != 0.f) g += a.weight_decay * p;
    if (a.momentum != 0.f) {
      m = first_step ? g : (a.momentum * m + (1.f - a.dampening) * g);
      g = a.nesterov ? (g + a.momentum * m) : m;
    }
    return p - a.lr * g;
  }
  if (a.kind == 2) p *= (1.f - a.lr * a.weight_decay);       // AdamW: decoupled decay
  else if (a.weight_decay != 0.f) g += a.weight_decay * p;    // Adam: L2
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v) * oc.bc2_inv_sqrt + a.eps;
  return p - a.lr * oc.bc1_inv * (m / denom);
}

ST_DEVINL int find_segment(const ParamSegment* segs, int n, int64_t e) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].offset <= e) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// mode 0: full step (exchange + update + publish); mode 1: publish only (build low-precision copies locally)
template <int MODE>
__global__ void __launch_bounds__(AR_THREADS) allreduce_update_kernel(const UpdateArgs a) {
  __shared__ ParamSegment s_segs[MAX_SEGS_SMEM];
  __shared__ uint32_t s_epoch;
  __shared__ int s_stop;          // 1: this iteration does not update (early stop now or latched)
  __shared__ int s_stop_now;
  __shared__ float s_loss;
  __shared__ int s_last;
  const int W = a.world, me = a.rank;
  const ParamSegment* segs = a.segs;
  if (a.n_segs <= MAX_SEGS_SMEM) {
    for (int i = threadIdx.x; i < a.n_segs; i += blockDim.x) s_segs[i] = a.segs[i];
    segs = s_segs;
  }
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    s_epoch = a.state->epoch + 1;
    s_stop = 0; s_stop_now = 0; s_loss = 0.f;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;

  if (MODE == 0 && W > 1) block_barrier_all_gpus<false>(a.flags, W, me, 0, epoch);

  // ---- owned shard (in float4 units; the arena is a multiple of 64 elements)
  const int64_t n4 = a.numel >> 2;
  int64_t lo4 = 0, hi4 = n4;
  if (MODE == 0 && W > 1) {
    const int64_t per = (n4 + W - 1) / W;
    lo4 = per * me;
    hi4 = lo4 + per < n4 ? lo4 + per : n4;
    if (lo4 > n4) lo4 = n4;
  }
  // Issue the first gradient loads (NVLink round trip ~2 us) BEFORE the early-stop decision below, so the
  // latency of the peer loads overlaps the tail-slot loads and the block-wide sync of that decision.
  const int64_t i4_first = lo4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  float4 g_first = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE == 0 && i4_first < hi4) {
    if (a.mc_grad != nullptr && W > 1) {
      g_first = ld_reduce_mc_f4(a.mc_grad + (i4_first << 2));
    } else {
      g_first = reinterpret_cast<const float4*>(a.grad[me])[i4_first];
      for (int r = 1; r < W; ++r) {
        const float4 q = ld_f4_stream(reinterpret_cast<const float4*>(a.grad[(me + r) % W]) + i4_first);
        g_first.x += q.x; g_first.y += q.y; g_first.z += q.z; g_first.w += q.w;
      }
    }
  }

  // ---- loss + early stopping: every block of every rank computes the same decision
  if (MODE == 0 && threadIdx.x == 0) {
    float loss = 0.f, votes = 0.f;
    for (int r = 0; r < W; ++r) {
      const float* tail = a.grad[r] + a.numel;
      loss += *reinterpret_cast<const volatile float*>(tail + 0);
      votes += *reinterpret_cast<const volatile float*>(tail + 1);
    }
    loss /= float(W);
    int stop_now = 0;
    const StepState st0 = *a.state;
    if (st0.stopped) {
      s_stop = 1;
    } else if (a.patience > 0) {
      if (!st0.es_seeded) {
      } else if (isnan(loss)) {
        stop_now = 1;
      } else {
        const uint32_t bad = (loss < st0.es_best) ? 0u : st0.es_bad + 1u;
        if (bad >= uint32_t(a.patience)) stop_now = 1;
      }
      if (votes > 0.f) stop_now = 1;
      s_stop = stop_now;
      s_stop_now = stop_now;
    }
    s_loss = loss;
  }
  __syncthreads();
  const bool skip_update = (MODE == 0) && (s_stop != 0);

  __shared__ OptConsts s_oc;
  __shared__ long long s_step1;
  if (threadIdx.x == 32 || (blockDim.x <= 32 && threadIdx.x == 0)) {   // one thread per block pays for the powf
    const int64_t st1 = a.state->step + 1;                             // 1-based step being applied
    s_step1 = st1;
    OptConsts t{1.f, 1.f};
    if (a.kind != 0) {
      t.bc1_inv = 1.f / (1.f - powf(a.beta1, float(st1)));
      t.bc2_inv_sqrt = rsqrtf(1.f - powf(a.beta2, float(st1)));
    }
    s_oc = t;
  }
  __syncthreads();
  const OptConsts oc = s_oc;
  const int64_t step1 = s_step1;
  const float inv_w = 1.f / float(W);
  const int n_pub = (MODE == 0) ? W : 1;

  if (!skip_update) {
    for (int64_t i4 = lo4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i4 < hi4; i4 += int64_t(gridDim.x) * blockDim.x) {
      const int64_t e = i4 << 2;
      float4 p = reinterpret_cast<const float4*>(a.param[me])[i4];
      if (MODE == 0) {
        float4 g;
        if (i4 == i4_first) {
          g = g_first;
        } else if (a.mc_grad != nullptr && W > 1) {
          g = ld_reduce_mc_f4(a.mc_grad + e);
        } else {
          g = reinterpret_cast<const float4*>(a.grad[me])[i4];
          for (int r = 1; r < W; ++r) {
            const int peer = (me + r) % W;      // stagger peers so links are used evenly
            const float4 q = ld_f4_stream(reinterpret_cast<const float4*>(a.grad[peer]) + i4);
            g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
          }
        }
        g.x *= inv_w; g.y *= inv_w; g.z *= inv_w; g.w *= inv_w;
        // optimizer state is indexed by the position inside the owned shard
        const int64_t s4 = i4 - lo4;
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
        const bool has_m = (a.kind != 0) || (a.momentum != 0.f);
        if (has_m) m = reinterpret_cast<const float4*>(a.m)[s4];
        if (a.kind != 0) v = reinterpret_cast<const float4*>(a.v)[s4];
        const bool first = (step1 == 1);
        p.x = opt_update_one(a, oc, p.x, g.x, m.x, v.x, first);
        p.y = opt_update_one(a, oc, p.y, g.y, m.y, v.y, first);
        p.z = opt_update_one(a, oc, p.z, g.z, m.z, v.z, first);
        p.w = opt_update_one(a, oc, p.w, g.w, m.w, v.w, first);
        if (has_m) reinterpret_cast<float4*>(a.m)[s4] = m;
        if (a.kind != 0) reinterpret_cast<float4*>(a.v)[s4] = v;
      }
      // ---- publish
      const ParamSegment sg = segs[find_segment(segs, a.n_segs, e)];
      const int64_t local = e - sg.offset;            // index inside the parameter (may be padding)
      const int64_t count = int64_t(sg.rows) * sg.cols;
      const float pv[4] = {p.x, p.y, p.z, p.w};
      // NVLS publication: one multimem.st per value reaches every replica (this one included) through the switch - 1/W of the
      // egress of the per-peer loop below.  Used for the fp32 master copy and the packed row-major bf16 copy; transposed or
      // unaligned bf16 layouts keep the per-peer path.
      bool mc_done_f32 = false, mc_done_bf16 = false;
      if (MODE == 0 && W > 1 && a.mc_param != nullptr) {
        if (sg.publish_f32) { st_mc_f4(a.mc_param + e, p); mc_done_f32 = true; }
        if (a.mc_wbf16 != nullptr && sg.w_off >= 0 && sg.wt_off < 0 && (sg.cols & 3) == 0 && (sg.w_ld & 3) == 0 && (sg.w_off & 3) == 0 &&
            local + 3 < count) {
          const int row = int(local / sg.cols), col = int(local % sg.cols);
          const uint32_t lo = pack_bf16x2(pv[0], pv[1]), hi = pack_bf16x2(pv[2], pv[3]);
          asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(a.mc_wbf16 + sg.w_off + int64_t(row) * sg.w_ld + col),
                       "f"(__uint_as_float(lo)), "f"(__uint_as_float(hi)) : "memory");
          mc_done_bf16 = true;
        }
      }
      for (int k = 0; k < n_pub; ++k) {
        const int peer = (MODE == 0) ? (me + k) % W : me;
        if (mc_done_f32 && (mc_done_bf16 || (sg.w_off < 0 && sg.wt_off < 0)) && peer != me) continue;   // everything already sent
        if (MODE == 0 && ((sg.publish_f32 && !mc_done_f32) || peer == me)) reinterpret_cast<float4*>(a.param[peer])[i4] = p;
        if ((sg.w_off >= 0 || sg.wt_off >= 0) && !mc_done_bf16) {
          __nv_bfloat16* wb = a.wbf16[peer];
          const bool packed = sg.w_off >= 0 && sg.wt_off < 0 && (sg.cols & 3) == 0 && (sg.w_ld & 3) == 0 &&
                              (sg.w_off & 3) == 0 && local + 3 < count;
          if (packed) {   // 4 consecutive columns of one row -> one coalesced 8-byte store
            const int row = int(local / sg.cols), col = int(local % sg.cols);
            uint2 h;
            h.x = pack_bf16x2(pv[0], pv[1]);
            h.y = pack_bf16x2(pv[2], pv[3]);
            *reinterpret_cast<uint2*>(wb + sg.w_off + int64_t(row) * sg.w_ld + col) = h;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t li = local + j;
              if (li < count) {
                const int row = int(li / sg.cols), col = int(li % sg.cols);
                const __nv_bfloat16 h = __float2bfloat16(pv[j]);
                if (sg.w_off >= 0) wb[sg.w_off + int64_t(row) * sg.w_ld + col] = h;
                if (sg.wt_off >= 0) wb[sg.wt_off + int64_t(col) * sg.wt_ld + row] = h;
              }
            }
          }
        }
      }
    }
  }

  if (MODE == 0 && W > 1) block_barrier_all_gpus<true>(a.flags, W, me, 1, epoch);

  if (MODE == 0 && a.zero_grads) {
    // Clear the local gradient arena for the next step.  Block b may only clear what
    // the b-th block of every rank has finished reading (that is what barrier B of
    // block b guarantees): the b-th stride of every rank's shard.  The tail slots are
    // read by all blocks of all ranks and are cleared by the last block below.
    float4* g4 = reinterpret_cast<float4*>(a.grad[me]);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t per = (W > 1) ? (n4 + W - 1) / W : n4;
    for (int r = 0; r < W; ++r) {
      const int64_t rlo = per * r < n4 ? per * r : n4;
      const int64_t rhi = rlo + per < n4 ? rlo + per : n4;
      for (int64_t i = rlo + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < rhi; i += int64_t(gridDim.x) * blockDim.x) g4[i] = z;
    }
  }

  // ---- last block to finish advances the step state
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(&a.state->blocks_done, 1u);
    s_last = (done == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    StepState* st = a.state;
    st->blocks_done = 0;
    if (MODE == 0) {
      st->epoch = epoch;
      if (a.zero_grads)
        for (int k = 0; k < a.tail; ++k) a.grad[me][a.numel + k] = 0.f;
      const int64_t it = st->iter;
      if (!st->stopped) {
        if (a.loss_hist != nullptr && it < a.max_iters) a.loss_hist[it] = s_loss;
        st->last_loss = s_loss;
        if (a.patience > 0) {
          if (!st->es_seeded) { st->es_seeded = 1; st->es_best = s_loss; st->es_bad = 0; }
          else if (!isnan(s_loss)) {
            if (s_loss < st->es_best) { st->es_best = s_loss; st->es_bad = 0; } else { st->es_bad += 1; }
          }
        }
        if (s_stop_now) st->stopped = 1; else st->step += 1;
        st->iter = it + 1;
        if (a.host_mailbox != nullptr) {
          a.host_mailbox[2] = __float_as_uint(s_loss);
          a.host_mailbox[1] = uint32_t(it + 1);
          a.host_mailbox[0] = st->stopped;
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- one-shot push variant
// For small arenas the exchange is latency bound, so the data movement is taken off the critical path entirely:
// the weight-gradient GEMMs (and the bias-gradient / loss reductions) mirror every write of the gradient arena into
// "their" receive slot on every peer while they compute (GemmEpilogue::push_*).  When this kernel starts, all it
// needs is ONE flag exchange ("my pushes are complete"); after that every operand of the reduction is LOCAL:
//     g = (local + sum_p recv[parity][p]) / W  ->  Adam/SGD on the full replica (m, v replicated)  ->  fp32 + bf16
// No second barrier: receive slots are double buffered by iteration parity - a peer can only write slot parity
// (t+2)&1 after it has seen my flag of step t+1, which I raise after this kernel (incl. its clearing of the slots)
// has finished.
__global__ void __launch_bounds__(AR_THREADS) allreduce_push_update_kernel(const UpdateArgs a) {
  __shared__ ParamSegment s_segs[MAX_SEGS_SMEM];
  __shared__ uint32_t s_epoch;
  __shared__ int s_stop, s_stop_now, s_last;
  __shared__ float s_loss;
  __shared__ OptConsts s_oc;
  __shared__ long long s_step1;
  const int W = a.world, me = a.rank;
  const ParamSegment* segs = a.segs;
  if (a.n_segs <= MAX_SEGS_SMEM) {
    for (int i = threadIdx.x; i < a.n_segs; i += blockDim.x) s_segs[i] = a.segs[i];
    segs = s_segs;
  }
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    s_epoch = a.state->epoch + 1;
    s_stop = 0; s_stop_now = 0; s_loss = 0.f;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int par = int(a.state->iter & 1);
  // "my pushes are complete" (they were issued by earlier kernels of this stream; release is cumulative over them)
  if (W > 1) block_barrier_all_gpus<true>(a.flags, W, me, 0, epoch);

  const float* rbase = a.recv_local + int64_t(par) * W * a.recv_stride;
  if (threadIdx.x == 0) {
    float loss = 0.f, votes = 0.f;
    for (int p = 0; p < W; ++p) {
      const float* tp = (p == me) ? (a.grad[me] + a.numel) : (rbase + p * a.recv_stride + a.numel);
      loss += tp[0];
      votes += tp[1];
    }
    loss /= float(W);
    int stop_now = 0;
    const StepState st0 = *a.state;
    if (st0.stopped) {
      s_stop = 1;
    } else if (a.patience > 0) {
      if (!st0.es_seeded) {
      } else if (isnan(loss)) {
        stop_now = 1;
      } else {
        const uint32_t bad = (loss < st0.es_best) ? 0u : st0.es_bad + 1u;
        if (bad >= uint32_t(a.patience)) stop_now = 1;
      }
      if (votes > 0.f) stop_now = 1;
      s_stop = stop_now; s_stop_now = stop_now;
    }
    s_loss = loss;
    const int64_t st1 = st0.step + 1;
    s_step1 = st1;
    OptConsts t{1.f, 1.f};
    if (a.kind != 0) {
      t.bc1_inv = 1.f / (1.f - powf(a.beta1, float(st1)));
      t.bc2_inv_sqrt = rsqrtf(1.f - powf(a.beta2, float(st1)));
    }
    s_oc = t;
  }
  __syncthreads();
  const bool skip_update = s_stop != 0;
  const OptConsts oc = s_oc;
  const bool first = s_step1 == 1;
  const float inv_w = 1.f / float(W);
  const int64_t n4 = a.numel >> 2;
  const bool has_m = (a.kind != 0) || (a.momentum != 0.f);
  float4* G = reinterpret_cast<float4*>(a.grad[me]);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i4 < n4; i4 += int64_t(gridDim.x) * blockDim.x) {
    // canonical summation order (rank 0 .. W-1, the local arena standing in for slot `me`): every replica adds
    // the same numbers in the same order, so all ranks compute bit-identical parameters
    float4 g = z4;
    for (int p = 0; p < W; ++p) {
      float4* rp = (p == me) ? (G + i4) : (reinterpret_cast<float4*>(const_cast<float*>(rbase) + p * a.recv_stride) + i4);
      const float4 q = *rp;
      g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      *rp = z4;                                  // cleared for the next iteration of this parity
    }
    if (skip_update) continue;
    g.x *= inv_w; g.y *= inv_w; g.z *= inv_w; g.w *= inv_w;
    float4 p = reinterpret_cast<const float4*>(a.param[me])[i4];
    float4 m = z4, v = z4;
    if (has_m) m = reinterpret_cast<const float4*>(a.m)[i4];
    if (a.kind != 0) v = reinterpret_cast<const float4*>(a.v)[i4];
    p.x = opt_update_one(a, oc, p.x, g.x, m.x, v.x, first);
    p.y = opt_update_one(a, oc, p.y, g.y, m.y, v.y, first);
    p.z = opt_update_one(a, oc, p.z, g.z, m.z, v.z, first);
    p.w = opt_update_one(a, oc, p.w, g.w, m.w, v.w, first);
    if (has_m) reinterpret_cast<float4*>(a.m)[i4] = m;
    if (a.kind != 0) reinterpret_cast<float4*>(a.v)[i4] = v;
    reinterpret_cast<float4*>(a.param[me])[i4] = p;
    const int64_t e = i4 << 2;
    const ParamSegment sg = segs[find_segment(segs, a.n_segs, e)];
    if (sg.w_off >= 0) {
      const int64_t local = e - sg.offset, count = int64_t(sg.rows) * sg.cols;
      __nv_bfloat16* wb = a.wbf16[me];
      const float pv[4] = {p.x, p.y, p.z, p.w};
      if ((sg.cols & 3) == 0 && (sg.w_ld & 3) == 0 && (sg.w_off & 3) == 0 && local + 3 < count) {
        const int row = int(local / sg.cols), col = int(local % sg.cols);
        uint2 h;
        h.x = pack_bf16x2(pv[0], pv[1]);
        h.y = pack_bf16x2(pv[2], pv[3]);
        *reinterpret_cast<uint2*>(wb + sg.w_off + int64_t(row) * sg.w_ld + col) = h;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (local + j < count) wb[sg.w_off + int64_t((local + j) / sg.cols) * sg.w_ld + int((local + j) % sg.cols)] = __float2bfloat16(pv[j]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&a.state->blocks_done, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    StepState* st = a.state;
    st->blocks_done = 0;
    st->epoch = epoch;
    for (int k = 0; k < a.tail; ++k) {
      a.grad[me][a.numel + k] = 0.f;
      for (int p = 0; p < W; ++p)
        if (p != me) const_cast<float*>(rbase)[p * a.recv_stride + a.numel + k] = 0.f;
    }
    const int64_t it = st->iter;
    if (!st->stopped) {
      if (a.loss_hist != nullptr && it < a.max_iters) a.loss_hist[it] = s_loss;
      st->last_loss = s_loss;
      if (a.patience > 0) {
        if (!st->es_seeded) { st->es_seeded = 1; st->es_best = s_loss; st->es_bad = 0; }
        else if (!isnan(s_loss)) {
          if (s_loss < st->es_best) { st->es_best = s_loss; st->es_bad = 0; } else { st->es_bad += 1; }
        }
      }
      if (s_stop_now) st->stopped = 1; else st->step += 1;
      st->iter = it + 1;
      if (a.host_mailbox != nullptr) {
        a.host_mailbox[2] = __float_as_uint(s_loss);
        a.host_mailbox[1] = uint32_t(it + 1);
        a.host_mailbox[0] = st->stopped;
      }
    }
  }
}

static int ar_grid(int64_t numel, int world) {
  const int64_t n4 = (numel >> 2);
  const int64_t per = (n4 + world - 1) / world;
  int64_t g = (per + AR_THREADS - 1) / AR_THREADS;
  if (g < 1) g = 1;
  if (g > AR_MAX_BLOCKS) g = AR_MAX_BLOCKS;
  return int(g);
}

int allreduce_update(const UpdateArgs& a, cudaStream_t stream) {
  if (a.world > 8 || (a.numel & 63)) return -1;
  return int(launch_kernel(allreduce_update_kernel<0>, dim3(ar_grid(a.numel, a.world)), dim3(AR_THREADS), 0, stream, a));
}

int allreduce_push_update(const UpdateArgs& a, cudaStream_t stream) {
  if (a.world > 8 || (a.numel & 63) || a.recv_local == nullptr) return -1;
  return int(launch_kernel(allreduce_push_update_kernel, dim3(ar_grid(a.numel, 1)), dim3(AR_THREADS), 0, stream, a));
}

int publish_params(const UpdateArgs& a, cudaStream_t stream) {
  if (a.numel & 63) return -1;
  return int(launch_kernel(allreduce_update_kernel<1>, dim3(ar_grid(a.numel, 1)), dim3(AR_THREADS), 0, stream, a));
}

// ----------------------------------------------------------------------------- plain all-reduce (bandwidth sweep, buffers)
struct SumArgs {
  float* bufs[8];
  uint32_t* flags[8];
  float* mc;
  StepState* state;
  int world, rank;
  int64_t n;
};

__global__ void __launch_bounds__(AR_THREADS) allreduce_sum_kernel(const SumArgs a) {
  __shared__ uint32_t s_epoch;
  __shared__ int s_last;
  if (threadIdx.x == 0) s_epoch = a.state->epoch + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int W = a.world, me = a.rank;
  block_barrier_all_gpus<false>(a.flags, W, me, 0, epoch);
  const int64_t n4 = a.n >> 2;
  const int64_t per = (n4 + W - 1) / W;
  const int64_t lo4 = per * me < n4 ? per * me : n4;
  const int64_t hi4 = lo4 + per < n4 ? lo4 + per : n4;
  // 4 independent 16-byte reductions in flight per thread (the loop is latency bound over NVLink otherwise)
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i4 = lo4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  constexpr int U = 4;
  for (; i4 + (U - 1) * stride < hi4; i4 += U * stride) {
    float4 g[U];
    if (a.mc != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u) g[u] = ld_reduce_mc_f4(a.mc + ((i4 + u * stride) << 2));
#pragma unroll
      for (int u = 0; u < U; ++u) st_mc_f4(a.mc + ((i4 + u * stride) << 2), g[u]);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) g[u] = reinterpret_cast<const float4*>(a.bufs[me])[i4 + u * stride];
      for (int r = 1; r < W; ++r) {
        const float4* src = reinterpret_cast<const float4*>(a.bufs[(me + r) % W]);
        float4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = ld_f4_stream(src + i4 + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) { g[u].x += q[u].x; g[u].y += q[u].y; g[u].z += q[u].z; g[u].w += q[u].w; }
      }
      for (int r = 0; r < W; ++r) {
        float4* dst = reinterpret_cast<float4*>(a.bufs[(me + r) % W]);
#pragma unroll
        for (int u = 0; u < U; ++u) dst[i4 + u * stride] = g[u];
      }
    }
  }
  for (; i4 < hi4; i4 += stride) {
    float4 g;
    if (a.mc != nullptr) {
      g = ld_reduce_mc_f4(a.mc + (i4 << 2));
      st_mc_f4(a.mc + (i4 << 2), g);
    } else {
      g = reinterpret_cast<const float4*>(a.bufs[me])[i4];
      for (int r = 1; r < W; ++r) {
        const float4 q = ld_f4_stream(reinterpret_cast<const float4*>(a.bufs[(me + r) % W]) + i4);
        g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      }
      for (int r = 0; r < W; ++r) reinterpret_cast<float4*>(a.bufs[(me + r) % W])[i4] = g;
    }
  }
  block_barrier_all_gpus<true>(a.flags, W, me, 1, epoch);
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&a.state->blocks_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    a.state->blocks_done = 0;
    a.state->epoch = epoch;
  }
}

int allreduce_sum_f32(float* const* bufs, uint32_t* const* flags, StepState* state, int world, int rank, int64_t n,
                      float* multicast, cudaStream_t stream) {
  if (world > 8 || (n & 3)) return -1;
  SumArgs a{};
  for (int r = 0; r < world; ++r) { a.bufs[r] = bufs[r]; a.flags[r] = flags[r]; }
  a.mc = multicast; a.state = state; a.world = world; a.rank = rank; a.n = n;
  const int64_t per4 = ((n >> 2) + world - 1) / world;
  int64_t g = (per4 + AR_THREADS - 1) / AR_THREADS;      // one 16-byte element per thread until the grid is full
  if (g < 1) g = 1;
  if (g > AR_MAX_BLOCKS) g = AR_MAX_BLOCKS;
  allreduce_sum_kernel<<<int(g), AR_THREADS, 0, stream>>>(a);
  return int(cudaGetLastError());
}

}  // namespace st
