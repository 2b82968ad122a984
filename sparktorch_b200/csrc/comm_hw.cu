// Hogwild (asynchronous parameter-server) training over NVLink peer memory.
//
// The reference runs a Flask process on the driver; every worker iteration GETs a
// dill'ed state_dict and POSTs dill'ed gradients over HTTP, the server assigns
// .grad and steps the optimizer under an optional global lock (hogwild.py:96-140,
// server.py:95-149; wire table SURVEY.md §2.4 X8-X10).
//
// Here the "server" is the memory of the GPUs themselves: the flat parameter arena
// and its optimizer state are sharded over the W GPUs (owner r holds shard r).
// There is no server process and no barrier between workers:
//
//   hw_pull : a worker copies every shard from its owner with 128-bit P2P loads
//             (the parameter all-gather; optionally also emitting the bf16 copies
//             the tensor-core GEMMs consume).
//   hw_push : the worker applies ITS gradient to every owner's shard remotely –
//             load p/m/v over NVLink, Adam/SGD update, store back.  One optimizer
//             step per push, exactly like the server's POST /update handler.
//
// acquireLock=True  -> lock striping: every (shard, stripe) has a spin lock in the
//                      owner's memory (system-scope CAS); pushes and pulls of a stripe
//                      are mutually exclusive, so no torn reads/updates.
// acquireLock=False -> lock-free hogwild: racing updates may be lost (by design).
//
// Early stopping follows the server's /losses handler: losses are accumulated in a
// window on rank 0; whoever completes a window (> window_len entries) steps the
// patience record and, on stop, raises the stop flag on every GPU.
#include "common.cuh"
#include "kernels.h"

namespace st {

constexpr int HW_THREADS = 256;

ST_DEVINL uint32_t atom_cas_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.acq_rel.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
ST_DEVINL uint32_t atom_add_sys(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.sys.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
ST_DEVINL float atom_add_sys_f32(float* p, float v) {
  float old;
  asm volatile("atom.relaxed.sys.global.add.f32 %0, [%1], %2;" : "=f"(old) : "l"(p), "f"(v) : "memory");
  return old;
}

ST_DEVINL void stripe_lock(uint32_t* lock) {
  if (threadIdx.x == 0) {
    while (atom_cas_sys(lock, 0u, 1u) != 0u) {
      __nanosleep(64);
    }
  }
  __syncthreads();
}
ST_DEVINL void stripe_unlock(uint32_t* lock) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys_u32(lock, 0u);
  }
}

// stripe geometry shared by pull and push: shard r = float4 range [per*r, per*(r+1)), cut into `stripes` pieces
struct StripeRange { int64_t lo4, hi4; };
ST_DEVINL StripeRange stripe_range(int64_t numel, int world, int stripes, int r, int s) {
  const int64_t n4 = numel >> 2;
  const int64_t per = (n4 + world - 1) / world;
  const int64_t slo = per * r < n4 ? per * r : n4;
  const int64_t shi = slo + per < n4 ? slo + per : n4;
  const int64_t sper = (per + stripes - 1) / stripes;
  StripeRange o;
  o.lo4 = slo + sper * s < shi ? slo + sper * s : shi;
  o.hi4 = o.lo4 + sper < shi ? o.lo4 + sper : shi;
  return o;
}

// grid = (stripes, world); block (s, k) serves shard (rank + k) % world so workers start on different owners
__global__ void __launch_bounds__(HW_THREADS) hw_pull_kernel(const HogwildArgs a) {
  const int r = (a.rank + blockIdx.y) % a.world;
  const int s = blockIdx.x;
  const StripeRange rg = stripe_range(a.numel, a.world, gridDim.x, r, s);
  uint32_t* lock = a.ctrl[r] + HW_CTRL_LOCKS + s;
  if (a.acquire_lock) stripe_lock(lock);
  const float4* src = reinterpret_cast<const float4*>(a.param[r]);
  float4* dst = reinterpret_cast<float4*>(a.local_param);
  for (int64_t i = rg.lo4 + threadIdx.x; i < rg.hi4; i += blockDim.x) dst[i] = ld_f4_stream(src + i);
  if (a.acquire_lock) stripe_unlock(lock);
}

__global__ void __launch_bounds__(HW_THREADS) hw_push_kernel(const HogwildArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  if (a.dry_run) return;
  const int r = (a.rank + blockIdx.y) % a.world;
  const int s = blockIdx.x;
  const StripeRange rg = stripe_range(a.numel, a.world, gridDim.x, r, s);
  uint32_t* ctrl = a.ctrl[r];
  uint32_t* lock = ctrl + HW_CTRL_LOCKS + s;
  __shared__ uint32_t s_step;
  if (a.acquire_lock) stripe_lock(lock);
  if (threadIdx.x == 0) s_step = atom_add_sys(ctrl + HW_CTRL_STEPS + s, 1u) + 1u;   // optimizer steps applied to this stripe
  __syncthreads();
  const uint32_t step1 = s_step;
  float bc1_inv = 1.f, bc2_inv_sqrt = 1.f;
  if (a.kind != 0) {
    bc1_inv = 1.f / (1.f - powf(a.beta1, float(step1)));
    bc2_inv_sqrt = rsqrtf(1.f - powf(a.beta2, float(step1)));
  }
  float4* P = reinterpret_cast<float4*>(a.param[r]);
  float4* Mv = reinterpret_cast<float4*>(a.m[r]);
  float4* Vv = reinterpret_cast<float4*>(a.v[r]);
  const float4* G = reinterpret_cast<const float4*>(a.local_grad);
  const bool has_m = (a.kind != 0) || (a.momentum != 0.f);
  for (int64_t i = rg.lo4 + threadIdx.x; i < rg.hi4; i += blockDim.x) {
    float4 p = ld_f4_stream(P + i);
    float4 g = G[i];
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
    if (has_m) m = ld_f4_stream(Mv + i);
    if (a.kind != 0) v = ld_f4_stream(Vv + i);
    float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = a.maximize ? -gp[j] : gp[j];
      if (a.kind == 0) {
        if (a.weight_decay != 0.f) gg += a.weight_decay * pp[j];
        if (a.momentum != 0.f) {
          mp[j] = (step1 == 1) ? gg : (a.momentum * mp[j] + (1.f - a.dampening) * gg);
          gg = a.nesterov ? (gg + a.momentum * mp[j]) : mp[j];
        }
        pp[j] -= a.lr * gg;
      } else {
        if (a.kind == 2) pp[j] *= (1.f - a.lr * a.weight_decay);
        else if (a.weight_decay != 0.f) gg += a.weight_decay * pp[j];
        mp[j] = a.beta1 * mp[j] + (1.f - a.beta1) * gg;
        vp[j] = a.beta2 * vp[j] + (1.f - a.beta2) * gg * gg;
        pp[j] -= a.lr * bc1_inv * (mp[j] / (sqrtf(vp[j]) * bc2_inv_sqrt + a.eps));
      }
    }
    P[i] = p;
    if (has_m) Mv[i] = m;
    if (a.kind != 0) Vv[i] = v;
    if (a.local_grad_rw != nullptr) reinterpret_cast<float4*>(a.local_grad_rw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (a.acquire_lock) stripe_unlock(lock);

  // ---- windowed early stopping on rank 0's record (one block per push does it)
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    uint32_t* c0 = a.ctrl[0];
    float* f0 = reinterpret_cast<float*>(c0);
    const float loss = a.use_val_loss ? a.loss_in[1] : a.loss_in[0];
    atom_add_sys(c0 + HW_CTRL_PUSHES, 1u);
    if (a.patience > 0 && ld_acquire_sys_u32(c0 + HW_CTRL_STOP) == 0u) {
      atom_add_sys_f32(f0 + HW_CTRL_WIN_SUM, loss);
      const uint32_t cnt = atom_add_sys(c0 + HW_CTRL_WIN_CNT, 1u) + 1u;
      if (cnt > uint32_t(a.window_len)) {
        while (atom_cas_sys(c0 + HW_CTRL_ES_LOCK, 0u, 1u) != 0u) __nanosleep(64);
        const uint32_t n = ld_acquire_sys_u32(c0 + HW_CTRL_WIN_CNT);
        if (n > uint32_t(a.window_len)) {         // still ours to close
          const float mean = *reinterpret_cast<volatile float*>(f0 + HW_CTRL_WIN_SUM) / float(n);
          *reinterpret_cast<volatile float*>(f0 + HW_CTRL_WIN_SUM) = 0.f;
          *reinterpret_cast<volatile uint32_t*>(c0 + HW_CTRL_WIN_CNT) = 0u;
          volatile uint32_t* seeded = c0 + HW_CTRL_ES_SEEDED;
          volatile float* best = f0 + HW_CTRL_ES_BEST;
          volatile uint32_t* bad = c0 + HW_CTRL_ES_BAD;
          bool stop = false;
          if (!*seeded) { *seeded = 1u; *best = mean; *bad = 0u; }
          else if (isnan(mean)) stop = true;
          else if (mean < *best) { *best = mean; *bad = 0u; }
          else { *bad = *bad + 1u; if (*bad >= uint32_t(a.patience)) stop = true; }
          if (stop)
            for (int w = 0; w < a.world; ++w) st_release_sys_u32(a.ctrl[w] + HW_CTRL_STOP, 1u);
        }
        __threadfence_system();
        st_release_sys_u32(c0 + HW_CTRL_ES_LOCK, 0u);
      }
    }
    const float train_loss = a.loss_in[0];
    int64_t it = 0;
    if (a.local_state != nullptr) {
      it = a.local_state->iter;
      a.local_state->iter = it + 1;
      a.local_state->step = it + 1;
    }
    if (a.loss_hist != nullptr && it < a.max_iters) a.loss_hist[it] = train_loss;
    if (a.loss_slot_rw != nullptr) { a.loss_slot_rw[0] = 0.f; a.loss_slot_rw[1] = 0.f; }
    if (a.host_mailbox != nullptr) {
      a.host_mailbox[2] = __float_as_uint(train_loss);
      a.host_mailbox[1] = uint32_t(it + 1);
      a.host_mailbox[0] = ld_acquire_sys_u32(a.ctrl[a.rank] + HW_CTRL_STOP);
    }
  }
}

static dim3 hw_grid(const HogwildArgs& a) { return dim3(unsigned(a.stripes), unsigned(a.world)); }

int hogwild_pull(const HogwildArgs& a, cudaStream_t stream) {
  if (a.world > 8 || (a.numel & 63) || a.stripes < 1 || a.stripes > HW_MAX_STRIPES) return -1;
  hw_pull_kernel<<<hw_grid(a), HW_THREADS, 0, stream>>>(a);
  return int(cudaGetLastError());
}

int hogwild_push(const HogwildArgs& a, cudaStream_t stream) {
  if (a.world > 8 || (a.numel & 63) || a.stripes < 1 || a.stripes > HW_MAX_STRIPES) return -1;
  return int(launch_kernel(hw_push_kernel, hw_grid(a), dim3(HW_THREADS), 0, stream, a));
}

}  // namespace st
