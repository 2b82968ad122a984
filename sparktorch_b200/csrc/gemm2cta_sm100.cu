// C[M,N] (bf16) = act(alpha * A[M,K] B[N,K]^T + bias) with tcgen05.mma.cta_group::2: a CTA PAIR (one 2-CTA cluster = the two SMs
// of a TPC) computes a 256 x 256 tile.  Why: with one CTA per 128 x 256 tile every SM writes (TMA) and reads (tensor core)
// 12 KB of shared memory per 135-cycle MMA step = 180 B/clk against a 128 B/clk shared-memory pipe - the single-CTA kernels
// of gemm_sm100.cu stop at ~0.75 of the tensor peak for exactly that reason (profiles/gemm_vs_cublas.md).  In a pair each SM
// stages its 128 rows of A and only HALF of the B tile (128 of the 256 columns); the tensor cores of both SMs read both
// halves, so the per-SM traffic drops to 8 KB written + 8 KB read per step (~120 B/clk) and L2 -> SM traffic drops by a third.
//
// Persistent: grid = #SMs (one pair per TPC), tiles are walked round-robin, the accumulator is double buffered in TMEM
// (2 x 256 columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//
//   warp 0      TMA producer (both CTAs; `.cta_group::2` loads complete on the LEADER's mbarrier)
//   warp 1      MMA issuer (leader CTA only); tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs
//   warps 2-9   epilogue: tcgen05.ld -> bias / ReLU -> bf16 -> swizzled shared-memory staging -> TMA store (full 128-byte lines,
//               ragged edges clipped by the tensor map); arrive on the leader's "TMEM free" barrier once TMEM has been read
// Tiles are walked in groups of 8 tile rows (all pairs of a wave then share ~8 A panels and ~9 B panels: the wave's operands
// fit the L2; a plain row-major walk streams all of A from HBM once per tile column).
//
// Replaces the cuBLAS GEMMs behind every Linear / Conv2d forward of the reference (torch ops called at
// /root/reference/sparktorch/distributed.py:150,176).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "gemm_sm100.h"
#include "kernels.h"

namespace st {

namespace {

constexpr int G2_BM = 128;              // rows of A per CTA (pair: 256)
constexpr int G2_BN = 256;              // columns of the pair's tile
constexpr int G2_BN_HALF = 128;         // rows of B staged per CTA
constexpr int G2_BK = 64;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;
constexpr int G2_B_BYTES = G2_BN_HALF * G2_BK * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;      // 32 KB
constexpr int G2_STAGES = 6;
constexpr int G2_THREADS = 320;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_MISC = 2048;
constexpr int G2_CSTAGE_BYTES = 32 * 128;                    // per epilogue warp: 32 rows x 64 bf16 columns
constexpr int G2_CSTAGE_OFF = G2_STAGES * G2_STAGE_BYTES;    // 1024-byte aligned
constexpr int G2_MISC_OFF = G2_CSTAGE_OFF + G2_EPI_WARPS * G2_CSTAGE_BYTES;
constexpr int G2_SMEM = G2_MISC_OFF + G2_MISC + 1024;
constexpr int G2_GROUP_M = 8;                                // tile rows per rasterisation group
constexpr uint32_t G2_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: "the same object in CTA 0 of the pair"

struct G2Shared {
  uint64_t full_bar[G2_STAGES];         // leader: both CTAs' TMA bytes of a stage have landed
  uint64_t empty_bar[G2_STAGES];        // per CTA: the MMAs that read this stage are done (multicast commit)
  uint64_t tmem_full_bar[2];            // per CTA: accumulator stage complete (multicast commit)
  uint64_t tmem_empty_bar[2];           // leader: both CTAs' epilogue warps have drained the accumulator stage
  uint32_t tmem_ptr;
  uint32_t error;
};

ST_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ST_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ST_DEVINL void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
ST_DEVINL void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
ST_DEVINL void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// both CTAs of the pair issue this for their own shared memory; the bytes are accounted on the leader's barrier
ST_DEVINL void tma_load_2d_2cta(void* smem_dst, const void* desc, uint64_t* bar, int32_t c_inner, int32_t c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & G2_PEER_MASK), "r"(c_inner), "r"(c_outer)
      : "memory");
}
ST_DEVINL void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive once on the barrier at this shared-memory offset in every CTA of `mask` when all MMAs issued so far are complete
ST_DEVINL void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
ST_DEVINL void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {      // arrive on the same barrier object in CTA `cta` of the cluster
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// a wait that cannot hang the GPU: ~2 s, then trap (a protocol bug in a cluster kernel otherwise blocks both SMs for good)
ST_DEVINL void g2_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fff) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ll) __trap();
    }
  }
}

ST_DEVINL void tma_store_2d(const void* desc, const void* smem_src, int32_t c_inner, int32_t c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
ST_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
ST_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
ST_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// tile t of the walk -> (tile row, tile column): groups of G2_GROUP_M tile rows, column-major inside a group
ST_DEVINL void g2_tile(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = G2_GROUP_M * tiles_n;
  const int gidx = t / per_group;
  const int first_m = gidx * G2_GROUP_M;
  const int rows = min(G2_GROUP_M, tiles_m - first_m);
  const int r = t - gidx * per_group;
  tm = first_m + r % rows;
  tn = r / rows;
}

struct G2Args {
  __nv_bfloat16* C;
  int64_t ldc;
  const float* bias;
  float alpha;
  int relu;
  int M, N, K;
  int tiles_m, tiles_n;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_c, const G2Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  G2Shared* sh = reinterpret_cast<G2Shared*>(smem + G2_MISC_OFF);
  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  const int n_tiles = g.tiles_m * g.tiles_n;
  const int num_kb = (g.K + G2_BK - 1) / G2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full_bar[a], 1);
      mbar_init(&sh->tmem_empty_bar[a], 2 * G2_EPI_WARPS);      // every epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(&sh->tmem_ptr, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();                      // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_ptr;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    uint32_t kbc = 0;
    for (int t = pair; t < n_tiles; t += n_pairs) {
      int tm, tn;
      g2_tile(t, g.tiles_m, g.tiles_n, tm, tn);
      const int m0 = tm * (2 * G2_BM) + int(rank) * G2_BM;
      const int n0 = tn * G2_BN + int(rank) * G2_BN_HALF;
      for (int i = 0; i < num_kb; ++i, ++kbc) {
        const uint32_t s = kbc % G2_STAGES;
        g2_wait(&sh->empty_bar[s], ((kbc / G2_STAGES) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + s * G2_STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx(&sh->full_bar[s], 2 * G2_STAGE_BYTES);     // this CTA's and the peer's bytes
          tma_load_2d_2cta(sa, &tmap_a, &sh->full_bar[s], i * G2_BK, m0);
          tma_load_2d_2cta(sa + G2_A_BYTES, &tmap_b, &sh->full_bar[s], i * G2_BK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = make_idesc_bf16(2 * G2_BM, G2_BN);
      uint32_t kbc = 0, tc = 0;
      for (int t = pair; t < n_tiles; t += n_pairs, ++tc) {
        const uint32_t as = tc & 1;
        g2_wait(&sh->tmem_empty_bar[as], ((tc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * G2_BN;
        for (int i = 0; i < num_kb; ++i, ++kbc) {
          const uint32_t s = kbc % G2_STAGES;
          g2_wait(&sh->full_bar[s], (kbc / G2_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_base + s * G2_STAGE_BYTES;
          const uint64_t da = make_kmajor_sw128_desc(sa);
          const uint64_t db = make_kmajor_sw128_desc(sa + G2_A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k) umma_f16_2cta(tacc, da + 2 * k, db + 2 * k, idesc, (i | k) != 0);
            umma_commit_2cta(&sh->empty_bar[s], 3);
            if (i == num_kb - 1) umma_commit_2cta(&sh->tmem_full_bar[as], 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, 8 warps: TMEM lane quadrant x column half) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* cst = smem + G2_CSTAGE_OFF + (warp - 2) * G2_CSTAGE_BYTES;     // this warp's staging tile: 32 rows x 128 B, 128B-swizzled
    uint32_t tc = 0;
    for (int t = pair; t < n_tiles; t += n_pairs, ++tc) {
      const uint32_t as = tc & 1;
      int tm, tn;
      g2_tile(t, g.tiles_m, g.tiles_n, tm, tn);
      const int row0 = tm * (2 * G2_BM) + int(rank) * G2_BM + q * 32;       // first row of this warp's 32
      const int colw = tn * G2_BN + half * (G2_BN / 2);                       // first of this warp's 128 columns
      g2_wait(&sh->tmem_full_bar[as], (tc >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t(q * 32) << 16) + as * G2_BN + half * (G2_BN / 2);
#pragma unroll 1
      for (int c = 0; c < G2_BN / 2; c += 64) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(tacc + c, r0);
        tmem_ld_32x32(tacc + c + 32, r1);
        tmem_ld_wait();
        if (c + 64 == G2_BN / 2) {            // last TMEM read of this tile: the MMA warp may start the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&sh->tmem_empty_bar[as], 0);
        }
        const int col0 = colw + c;
        tma_store_wait_read();                // the previous TMA store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int v = 0; v < 8; ++v) {         // eight 16-byte chunks (8 columns each) of this lane's row
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = 8 * v + j;
            float y = __uint_as_float(jj < 32 ? r0[jj] : r1[jj - 32]) * g.alpha;
            if (g.bias != nullptr && col0 + jj < g.N) y += __ldg(g.bias + col0 + jj);
            if (g.relu) y = fmaxf(y, 0.f);
            x[j] = y;
          }
          uint4 w;
          w.x = pack_bf16x2(x[0], x[1]); w.y = pack_bf16x2(x[2], x[3]); w.z = pack_bf16x2(x[4], x[5]); w.w = pack_bf16x2(x[6], x[7]);
          *reinterpret_cast<uint4*>(cst + lane * 128 + ((v ^ (lane & 7)) << 4)) = w;
        }
        fence_proxy_async();                  // generic-proxy stores -> visible to the TMA engine
        __syncwarp();
        if (lane == 0 && row0 < g.M && col0 < g.N) {
          tma_store_2d(&tmap_c, cst, col0, row0);     // rows >= M / columns >= N are clipped by the tensor map
          tma_store_commit();
        }
        __syncwarp();
      }
    }
    tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();                      // nobody frees TMEM / leaves while the peer may still read this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace

// Returns 0 on success, -2 when the shape is not handled here (caller falls back to the single-CTA kernels).
int gemm_bf16_2cta(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N, int K, const float* bias,
                   float alpha, int relu, cudaStream_t stream) {
  if (M < 256 || N < 256 || K < 64 || (lda % 8) || (ldb % 8) || (ldc % 8)) return -2;
  CUtensorMap ta, tb;
  if ((reinterpret_cast<uintptr_t>(C) & 15) != 0) return -2;
  if (make_tmap_bf16_2d(&ta, A, uint64_t(M), uint64_t(K), uint64_t(lda), G2_BM) != 0) return -3;
  if (make_tmap_bf16_2d(&tb, B, uint64_t(N), uint64_t(K), uint64_t(ldb), G2_BN_HALF) != 0) return -3;
  CUtensorMap tc;
  if (make_tmap_bf16_2d(&tc, C, uint64_t(M), uint64_t(N), uint64_t(ldc), 32) != 0) return -3;
  static int n_sm = 0;
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(gemm_bf16_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM) != cudaSuccess) return -4;
    attr_set = true;
  }
  G2Args g;
  g.C = reinterpret_cast<__nv_bfloat16*>(C);
  g.ldc = ldc;
  g.bias = bias;
  g.alpha = alpha;
  g.relu = relu;
  g.M = M; g.N = N; g.K = K;
  g.tiles_m = (M + 2 * G2_BM - 1) / (2 * G2_BM);
  g.tiles_n = (N + G2_BN - 1) / G2_BN;
  const int n_tiles = g.tiles_m * g.tiles_n;
  int pairs = n_sm / 2;
  if (pairs > n_tiles) pairs = n_tiles;
  gemm_bf16_2cta_kernel<<<2 * pairs, G2_THREADS, G2_SMEM, stream>>>(ta, tb, tc, g);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace st
