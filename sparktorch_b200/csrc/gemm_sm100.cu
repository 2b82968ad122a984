// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM),
// operands staged by TMA (SWIZZLE_128B), warp-specialised:
//
//     D[M,N] = alpha * sum_k A[M,K] * B[N,K]        (A and B both K-major)
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  : epilogue – tcgen05.ld the accumulator, apply the fused
//                 epilogue and store straight from registers.
//
// Fused epilogue (all optional, runtime selected): + bias[n], ReLU, multiply by
// (mask_src[m,n] > 0) (= ReLU backward), row-major bf16 store, TRANSPOSED bf16
// store (feeds the next K-major GEMM without a transpose kernel), row-major or
// transposed fp32 store / red.add (split-K, gradient arena), column sums over m
// (bias gradients) with one red.add per column.
//
// This covers every GEMM of Linear fwd / dgrad / wgrad (reference hot ops G1, G2,
// G3) – the reference runs cuBLAS sgemm + separate ATen
// bias/ReLU/threshold_backward/sum kernels for the same work.
#include "common.cuh"
#include "gemm_sm100.h"
#include "kernels.h"

namespace st {

bool g_use_pdl = false;
int g_persistent_mode = 1;   // 0 never, 1 when tiles > #SMs, 2 always (tests)
int g_pair_mode = 1;         // 0: never hand big plain products to the CTA-pair kernel (tests / comparisons)

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (STAGE_BYTES * 8 <= 200 * 1024) ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int BAR_BYTES = 4096;   // mbarriers + tmem slot (first 512 B), bias tile (256 floats), per-CTA column statistics (2 x 256 floats)
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /*align slack*/;
  static constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
};

// One 32x(CHUNK) slab of the accumulator (thread = row, registers = columns): fused epilogue math and stores.
// Shared by the single-tile and the persistent kernel.
template <int CHUNK>
ST_DEVINL void epilogue_chunk(const GemmEpilogue& ep, const uint32_t (&r)[32], const uint4 (&mk)[4], bool have_mask,
                              const __nv_bfloat16* mrow, bool have_bias, const float* s_bias, int c0, int row, bool row_ok,
                              int col0, int N, int lane, int lbl_off, float* s_sum = nullptr, float* s_sq = nullptr) {
  const bool vec_ok = (col0 + CHUNK <= N);   // leading dims are multiples of 8 elements, col0 multiple of 16
  float loss_part = 0.f, colsum_part = 0.f;     // what this thread added to the loss slot / its bias-gradient column
  float acc[32];
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) {
    float v = __uint_as_float(r[j]) * ep.alpha;
    if (have_bias) v += s_bias[c0 + j];
    if (ep.relu) v = fmaxf(v, 0.f);
    acc[j] = v;
  }
  if (have_mask) {
    if (row_ok && vec_ok) {
      // decode the prefetched mask registers with constant indices only (no address-of: keeps them in registers)
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
        if (col0 + j < N) acc[j] = (__bfloat162float(mrow[col0 + j]) > 0.f) ? acc[j] : 0.f;
    }
  }
  // ---- row-major fp32 of the plain result (logits for inference / validation)
  if (ep.out_f32 != nullptr && row_ok && ep.loss_kind != 0) {
    float* op = ep.out_f32 + size_t(row) * ep.ld_f32 + col0;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j)
      if (col0 + j < N) op[j] = acc[j];
  }
  // ---- fused loss (single n-tile, N <= CHUNK): the thread owns the whole output row.
  //      acc[] becomes dLoss/dOutput so the stores / column sums below emit dZ and db.
  if (ep.loss_kind == 1) {          // softmax cross entropy, mean over M rows
    const int y = row_ok ? ep.ce_labels[row + lbl_off] : 0;
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
    for (int j = 0; j < CHUNK; ++j) acc[j] = (row_ok && j < N) ? (acc[j] * inv - (j == y ? 1.f : 0.f)) * ep.loss_scale : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
    loss_part = li * ep.loss_scale;
    if (lane == 0) red_add_f32(ep.loss_slot, loss_part);
  } else if (ep.loss_kind == 2) {   // mean squared error over M x N
    float li = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      float dlt = 0.f;
      if (row_ok && j < N) dlt = acc[j] - ep.mse_target[size_t(row + lbl_off) * ep.ld_target + j];
      li += dlt * dlt;
      acc[j] = 2.f * dlt * ep.loss_scale;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
    loss_part = li * ep.loss_scale;
    if (lane == 0) red_add_f32(ep.loss_slot, loss_part);
  }
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) acc[j] = 0.f;
  }
  // ---- row-major bf16
  if (ep.out_bf16 != nullptr && row_ok) {
    __nv_bfloat16* op = ep.out_bf16 + size_t(row) * ep.ld_out + col0;
    if (vec_ok || (col0 + CHUNK <= ep.ld_out && ep.loss_kind != 0)) {
      // (fused loss: the padding columns up to the pitch are written as zeros)
#pragma unroll
      for (int v = 0; v < CHUNK / 8; ++v) {
        uint4 w;
        w.x = pack_bf16x2(acc[8 * v + 0], acc[8 * v + 1]);
        w.y = pack_bf16x2(acc[8 * v + 2], acc[8 * v + 3]);
        w.z = pack_bf16x2(acc[8 * v + 4], acc[8 * v + 5]);
        w.w = pack_bf16x2(acc[8 * v + 6], acc[8 * v + 7]);
        if (col0 + 8 * v < ep.ld_out) reinterpret_cast<uint4*>(op)[v] = w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (col0 + j < N) op[j] = __float2bfloat16(acc[j]);
    }
  }
  // ---- row-major fp32 (store or red.add for split-K)
  if (ep.out_f32 != nullptr && row_ok && ep.loss_kind == 0) {
    float* op = ep.out_f32 + size_t(row) * ep.ld_f32 + col0;
    if (ep.atomic_f32) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (col0 + j < N) red_add_f32(op + j, acc[j]);
    } else if (vec_ok && (ep.ld_f32 & 3) == 0) {
#pragma unroll
      for (int v = 0; v < CHUNK / 4; ++v)
        reinterpret_cast<float4*>(op)[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (col0 + j < N) op[j] = acc[j];
    }
  }
  // ---- transposed stores: lanes hold consecutive rows -> coalesced along m
  if (ep.out_t_bf16 != nullptr) {
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      if (col0 + j < N && row_ok) ep.out_t_bf16[size_t(col0 + j) * ep.ld_out_t + row] = __float2bfloat16(acc[j]);
    }
  }
  if (ep.out_t_f32 != nullptr) {
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      if (col0 + j < N && row_ok) {
        float* op = ep.out_t_f32 + size_t(col0 + j) * ep.ld_f32_t + row;
        if (ep.atomic_f32) red_add_f32(op, acc[j]); else *op = acc[j];
      }
    }
  }
  // ---- column sums over the 32 rows of this warp (bias gradient)
  if (ep.colsum != nullptr) {
    // butterfly: after the 5 rounds lane L holds sum over lanes of column L (31 shuffles per 32 columns)
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
    colsum_part = v[0];
    if (lane < CHUNK && col0 + lane < N) {
      if (s_sum != nullptr) atomicAdd(s_sum + c0 + lane, colsum_part);     // persistent CTA: flushed once per column tile
      else red_add_f32(ep.colsum + col0 + lane, colsum_part);
    }
  }
  // ---- column sums of squares (with colsum: the batch-norm statistics of a convolution output, no extra pass)
  if (ep.colsumsq != nullptr) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (j < CHUNK) ? acc[j] * acc[j] : 0.f;
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
    if (lane < CHUNK && col0 + lane < N) {
      if (s_sq != nullptr) atomicAdd(s_sq + c0 + lane, v[0]);
      else red_add_f32(ep.colsumsq + col0 + lane, v[0]);
    }
  }
  // ---- GEMM -> all-gather push (one warp-uniform branch; nothing of this is issued on the normal path):
  //      repeat every write that went into the LOCAL gradient arena into my receive slot on each peer.
  if (ep.push_world > 1 && ep.push_mask != 0) {
    const int par = int(ep.push_state->iter & 1);
    for (int p = 0; p < ep.push_world; ++p) {
      if (p == ep.push_me) continue;
      const long long delta = ep.push_delta[par * ep.push_world + p];
      if ((ep.push_mask & 1) && ep.out_f32 != nullptr && row_ok && ep.loss_kind == 0) {
        float* op = reinterpret_cast<float*>(reinterpret_cast<char*>(ep.out_f32 + size_t(row) * ep.ld_f32 + col0) + delta);
        if (ep.atomic_f32) {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j)
            if (col0 + j < N) red_add_f32(op + j, acc[j]);
        } else if (vec_ok && (ep.ld_f32 & 3) == 0) {
#pragma unroll
          for (int v = 0; v < CHUNK / 4; ++v)
            reinterpret_cast<float4*>(op)[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j)
            if (col0 + j < N) op[j] = acc[j];
        }
      }
      if ((ep.push_mask & 1) && ep.out_t_f32 != nullptr && row_ok) {
#pragma unroll
        for (int j = 0; j < CHUNK; ++j)
          if (col0 + j < N) {
            float* op = reinterpret_cast<float*>(reinterpret_cast<char*>(ep.out_t_f32 + size_t(col0 + j) * ep.ld_f32_t + row) + delta);
            if (ep.atomic_f32) red_add_f32(op, acc[j]); else *op = acc[j];
          }
      }
      if ((ep.push_mask & 2) && ep.colsum != nullptr && lane < CHUNK && col0 + lane < N)
        red_add_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(ep.colsum + col0 + lane) + delta), colsum_part);
      if ((ep.push_mask & 4) && ep.loss_kind != 0 && lane == 0)
        red_add_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(ep.loss_slot) + delta), loss_part);
    }
  }
}

// A_MN / B_MN: the operand is stored "MN-major" = [K rows, M|N contiguous] (e.g. the
// activations [batch, features] when the reduction runs over the batch, or a weight
// [out, in] used as the B operand of dgrad).  TMA then brings 64-wide (128 B) column
// slabs of BLOCK_K rows; the UMMA descriptor walks them with LBO = slab size and
// SBO = 1024 B (8 K-rows), and each UMMA_K=16 step advances by 16 rows = 2048 B.
template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmEpilogue ep, const GemmWindow win, const GemmPeerB peer, int M, int N, int K,
                    int k_blocks_per_split) {
  using S = GemmSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full_bar = empty_bar + S::STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + 512);   // [BLOCK_N] (BAR_BYTES = 2048)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* prof = ep.prof ? ep.prof + 8 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  if (prof && threadIdx.x == 0) prof[0] = clock64();
  const int m0 = blockIdx.y * BLOCK_M;
  const int n0 = blockIdx.x * BLOCK_N;
  const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * k_blocks_per_split;
  const int kb_end = min(total_kb, kb_begin + k_blocks_per_split);
  const int num_kb = kb_end - kb_begin;   // >= 1 by construction of the grid
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], (!B_MN && peer.world > 0) ? 5 : 1);   // TMA (A) [+ 4 gather warps (B)]
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (prof && threadIdx.x == 0) prof[1] = clock64();
  // Everything above touched only on-chip state and kernel parameters; from here on we read what
  // earlier kernels of the step produced.
  pdl_wait();
  pdl_launch_dependents();
  if (prof && threadIdx.x == 0) prof[2] = clock64();
  // Device-side mini-batch window: the batch of iteration i is rows [off, off + window_rows) of the
  // resident partition; computed from the iteration counter the update kernel maintains, so the
  // captured graph walks through the data without any host involvement or gather kernel.
  int win_off = 0;
  if (win.state != nullptr) win_off = int((win.state->iter % win.n_windows) * win.window_rows);
  const int a_row_off = (win.flags & 1) ? win_off : 0;   // A K-major: outer (row) coordinate
  const int b_k_off = (win.flags & 2) ? win_off : 0;     // B MN-major: K (row) coordinate
  const int lbl_off = (win.flags & 4) ? win_off : 0;     // labels / targets of the fused loss


  // The two issuing warps run their loops warp-UNIFORMLY (every lane waits on the mbarrier, `elect.sync` picks the lane that
  // issues): from a `lane == 0` branch ptxas cannot keep the descriptors / addresses in uniform registers and wraps every
  // UTMALDG / UTCHMMA / UTCBAR in an R2UR waterfall loop (found with the K-STEP kernel, profiles/kstep_v1_timeline_n1.md).
  const uint32_t smem_base = smem_u32(smem);
  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % S::STAGES;
      const uint32_t ph = (i / S::STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + s * S::STAGE_BYTES;
        uint8_t* sb = sa + S::A_BYTES;
        const bool peer_b = !B_MN && peer.world > 0;
        mbar_arrive_expect_tx(&full_bar[s], peer_b ? S::A_BYTES : S::STAGE_BYTES);
        const int k0 = (kb_begin + i) * BLOCK_K;
        if constexpr (A_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + 64 * j, k0);
        } else {
          if (win.flags & 8) {       // implicit-GEMM convolution: one filter tap x 64 channels of the tile's 128 output pixels
            const int kb = kb_begin + i;
            const int tap = kb / win.conv_cb, cb = kb - tap * win.conv_cb;
            const int fr = tap / win.conv_S, fs = tap - fr * win.conv_S;
            const int img = m0 / (win.conv_P * win.conv_Q), rem = m0 - img * (win.conv_P * win.conv_Q);
            const int op = rem / win.conv_Q, oq = rem - op * win.conv_Q;
            tma_load_im2col_4d(sa, &tmap_a, &full_bar[s], cb * BLOCK_K, oq * win.conv_stride - win.conv_pad, op * win.conv_stride - win.conv_pad,
                               img, uint16_t(fs * win.conv_dil), uint16_t(fr * win.conv_dil));
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[s], k0, m0 + a_row_off);
          }
        }
        if constexpr (B_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + 64 * j, k0 + b_k_off);
        } else if (!peer_b) {
          tma_load_2d(sb, &tmap_b, &full_bar[s], k0, n0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N < 16 ? 16 : BLOCK_N) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % S::STAGES;
      const uint32_t ph = (i / S::STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_base + s * S::STAGE_BYTES;
      const uint32_t sb = sa + S::A_BYTES;
      const uint64_t da = A_MN ? make_mnmajor_sw128_desc(sa, BLOCK_K * 128) : make_kmajor_sw128_desc(sa);
      const uint64_t db = B_MN ? make_mnmajor_sw128_desc(sb, BLOCK_K * 128) : make_kmajor_sw128_desc(sb);
      // per UMMA_K=16 step, in units of 16 B of the start-address field:
      //   K-major : 16 bf16 = 32 B inside the swizzled 128 B row -> +2
      //   MN-major: 16 K-rows of 128 B = 2048 B                  -> +128
      constexpr uint64_t ka = A_MN ? 128 : 2, kb = B_MN ? 128 : 2;
      if (elect_one()) {
        if (prof && i == 0) prof[3] = clock64();
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          umma_f16(tmem_base, da + ka * k, db + kb * k, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[s]);            // frees the smem slot when these MMAs retire
        if (i == num_kb - 1) umma_commit(tmem_full_bar);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < M;
    constexpr int CHUNK = BLOCK_N >= 32 ? 32 : 16;
    // Everything that does not depend on the accumulator is fetched while the main loop runs:
    // the bias slice goes to shared memory, the first ReLU-mask chunk to registers.
    const int et = threadIdx.x - 64;           // 0..127
    const bool peer_b = !B_MN && peer.world > 0;
    // with split-K every split adds into the same output: only split 0 contributes the bias
    const bool have_bias = (ep.bias != nullptr || (peer_b && peer.bias_off >= 0)) && blockIdx.z == 0;
    if (peer_b && peer.bias_off >= 0) {
      for (int c = et; c < BLOCK_N; c += 128) {
        const int64_t f = peer.bias_off + n0 + c;
        s_bias[c] = (n0 + c < N) ? peer.param[f / peer.shard_elems][f] : 0.f;
      }
    } else if (ep.bias != nullptr) {
      for (int c = et; c < BLOCK_N; c += 128) s_bias[c] = (n0 + c < N) ? ep.bias[n0 + c] : 0.f;
    }
    const bool have_mask = ep.mask_src != nullptr;
    const __nv_bfloat16* mrow = have_mask ? ep.mask_src + size_t(row_ok ? row : 0) * ep.ld_mask : nullptr;
    uint4 mk[4], mk_next[4];
    auto load_mask = [&](uint4 (&dst)[4], int col0) {
      if (have_mask && row_ok && col0 + CHUNK <= N) {
        const uint4* mp = reinterpret_cast<const uint4*>(mrow + col0);
#pragma unroll
        for (int v = 0; v < CHUNK / 8; ++v) dst[v] = mp[v];
      }
    };
    load_mask(mk_next, n0);
    asm volatile("bar.sync 1, 128;" ::: "memory");   // s_bias visible to the four epilogue warps
    if constexpr (!B_MN && BLOCK_N <= 32) {
      if (peer_b) {
        // ===== K-AG producer: gather this CTA's B tiles from the parameter owners over NVLink =====
        // item = (row n, 16-byte chunk c) of a [BLOCK_N x 64] bf16 tile = 8 fp32 source values.
        constexpr int ITEMS = (BLOCK_N * 8 + 127) / 128;     // per thread per k-block (1 or 2)
        constexpr int DEPTH = ITEMS == 1 ? 8 : 4;            // k-blocks of loads in flight (hides ~2 us peer latency)
        float4 ring[DEPTH][ITEMS][2];
        const bool k_vec = (K & 7) == 0 && (peer.w_off & 3) == 0;
        auto issue = [&](int i, float4 (&dst)[ITEMS][2]) {
          const int k0 = (kb_begin + i) * BLOCK_K;
#pragma unroll
          for (int it = 0; it < ITEMS; ++it) {
            const int item = et + it * 128;
            const int n = item >> 3, c = item & 7;
            const int gn = n0 + n, gk = k0 + c * 8;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            if (item < BLOCK_N * 8 && gn < N && gk < K) {
              const int64_t f = peer.w_off + int64_t(gn) * K + gk;
              if (k_vec && gk + 8 <= K) {
                lo = ld_f4_stream(reinterpret_cast<const float4*>(peer.param[f / peer.shard_elems] + f));
                hi = ld_f4_stream(reinterpret_cast<const float4*>(peer.param[(f + 4) / peer.shard_elems] + f + 4));
              } else {
                float t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] = (gk + q < K) ? peer.param[(f + q) / peer.shard_elems][f + q] : 0.f;
                lo = make_float4(t[0], t[1], t[2], t[3]);
                hi = make_float4(t[4], t[5], t[6], t[7]);
              }
            }
            dst[it][0] = lo;
            dst[it][1] = hi;
          }
        };
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
          if (d < num_kb) issue(d, ring[d]);
        for (int i = 0; i < num_kb; ++i) {
          const int s = i % S::STAGES;
          const uint32_t ph = (i / S::STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sb = smem + s * S::STAGE_BYTES + S::A_BYTES;
          const int k0 = (kb_begin + i) * BLOCK_K;
#pragma unroll
          for (int d = 0; d < DEPTH; ++d) {
            if (d == i % DEPTH) {
#pragma unroll
              for (int it = 0; it < ITEMS; ++it) {
                const int item = et + it * 128;
                if (item < BLOCK_N * 8) {
                  const int n = item >> 3, c = item & 7;
                  const float4 lo = ring[d][it][0], hi = ring[d][it][1];
                  uint4 w;
                  w.x = pack_bf16x2(lo.x, lo.y); w.y = pack_bf16x2(lo.z, lo.w);
                  w.z = pack_bf16x2(hi.x, hi.y); w.w = pack_bf16x2(hi.z, hi.w);
                  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
                  *reinterpret_cast<uint4*>(sb + n * 128 + ((c ^ (n & 7)) << 4)) = w;
                  const int gn = n0 + n, gk = k0 + c * 8;
                  if (peer.wb_local != nullptr && blockIdx.y == 0 && gn < N && gk < peer.ld_wb)
                    *reinterpret_cast<uint4*>(peer.wb_local + int64_t(gn) * peer.ld_wb + gk) = w;
                }
              }
              if (i + DEPTH < num_kb) issue(i + DEPTH, ring[d]);
            }
          }
          fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[s]);
        }
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (prof && et == 0) prof[4] = clock64();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += CHUNK) {
      const int col0 = n0 + c0;
      if (col0 >= N) break;                    // warp-uniform
      uint32_t r[32];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c0);
      if constexpr (CHUNK == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
#pragma unroll
      for (int v = 0; v < 4; ++v) mk[v] = mk_next[v];
      if (col0 + CHUNK < N) load_mask(mk_next, col0 + CHUNK);   // prefetch the next chunk's mask
      tmem_ld_wait();
      epilogue_chunk<CHUNK>(ep, r, mk, have_mask, mrow, have_bias, s_bias, c0, row, row_ok, col0, N, lane, lbl_off);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) prof[5] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
  if (prof && threadIdx.x == 32) prof[6] = clock64();
}


// ----------------------------------------------------------------------------- persistent variant
// One CTA per SM loops over output tiles (m fastest, so concurrently running CTAs share the B tile in L2).
// The accumulator is double buffered in TMEM (2 x BLOCK_N columns): while the epilogue warps drain tile t,
// the MMA warp already accumulates tile t+1, and the TMA producer runs ahead across tile boundaries through
// the same smem ring.  Used for large problems (tiles > #SMs): big-batch Linear layers and conv GEMMs.
template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                            const GemmEpilogue ep, int M, int N, int K) {
  using S = GemmSmem<BLOCK_N>;
  static_assert(BLOCK_N >= 64 && 2 * BLOCK_N <= 512, "persistent tiles are 64..256 wide");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full_bar = empty_bar + S::STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);              // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch_dependents();

  const uint32_t smem_base = smem_u32(smem);
  if (warp == 0) {
    // whole warp runs the loop, one elected lane issues (see the single-tile kernel)
    uint32_t it = 0;                                // k-block counter across tiles (smem ring position)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile % tiles_m) * BLOCK_M, n0 = (tile / tiles_m) * BLOCK_N;
      for (int i = 0; i < num_kb; ++i, ++it) {
        const int s = it % S::STAGES;
        const uint32_t ph = (it / S::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + s * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
          const int k0 = i * BLOCK_K;
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + 64 * j, k0);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[s], k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + 64 * j, k0);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[s], k0, n0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
    constexpr uint64_t ka = A_MN ? 128 : 2, kb = B_MN ? 128 : 2;
    uint32_t it = 0, t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const uint32_t as = t & 1;
      mbar_wait(&tmem_empty_bar[as], ((t >> 1) & 1) ^ 1);      // epilogue drained this accumulator stage
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * BLOCK_N;
      for (int i = 0; i < num_kb; ++i, ++it) {
        const int s = it % S::STAGES;
        const uint32_t ph = (it / S::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        const uint64_t da = A_MN ? make_mnmajor_sw128_desc(sa, BLOCK_K * 128) : make_kmajor_sw128_desc(sa);
        const uint64_t db = B_MN ? make_mnmajor_sw128_desc(sb, BLOCK_K * 128) : make_kmajor_sw128_desc(sb);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_f16(tacc, da + ka * k, db + kb * k, idesc, (i | k) != 0);
          umma_commit(&empty_bar[s]);
          if (i == num_kb - 1) umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    constexpr int CHUNK = 32;
    const bool have_bias = ep.bias != nullptr;
    const bool have_mask = ep.mask_src != nullptr;
    // Batch-norm statistics (column sums + sums of squares): thousands of row tiles would hammer the same few cache
    // lines with global atomics, so a CTA accumulates in shared memory and flushes when it moves to another column tile.
    const bool stat_mode = ep.colsum != nullptr && ep.colsumsq != nullptr && ep.push_world <= 1;
    float* s_sum = stat_mode ? s_bias + 256 : nullptr;
    float* s_sq = stat_mode ? s_bias + 512 : nullptr;
    int stat_n0 = -1;
    auto flush_stats = [&]() {                                  // callers bracket this with bar.sync 1
      for (int c = et; c < BLOCK_N; c += 128) {
        if (stat_n0 >= 0 && stat_n0 + c < N) {
          red_add_f32(ep.colsum + stat_n0 + c, s_sum[c]);
          red_add_f32(ep.colsumsq + stat_n0 + c, s_sq[c]);
        }
        s_sum[c] = 0.f;
        s_sq[c] = 0.f;
      }
    };
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int m0 = (tile % tiles_m) * BLOCK_M, n0 = (tile / tiles_m) * BLOCK_N;
      const uint32_t as = t & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      asm volatile("bar.sync 1, 128;" ::: "memory");          // previous tile's s_bias readers / statistic adders are done
      if (stat_mode && n0 != stat_n0) {
        flush_stats();
        stat_n0 = n0;
      }
      if (have_bias)
        for (int c = et; c < BLOCK_N; c += 128) s_bias[c] = (n0 + c < N) ? ep.bias[n0 + c] : 0.f;
      const __nv_bfloat16* mrow = have_mask ? ep.mask_src + size_t(row_ok ? row : 0) * ep.ld_mask : nullptr;
      uint4 mk[4], mk_next[4];
      auto load_mask = [&](uint4 (&dst)[4], int col0) {
        if (have_mask && row_ok && col0 + CHUNK <= N) {
          const uint4* mp = reinterpret_cast<const uint4*>(mrow + col0);
#pragma unroll
          for (int v = 0; v < CHUNK / 8; ++v) dst[v] = mp[v];
        }
      };
      load_mask(mk_next, n0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tmem_full_bar[as], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += CHUNK) {
        const int col0 = n0 + c0;
        const bool live = col0 < N;                            // warp-uniform
        uint32_t r[32];
        if (live) {
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * BLOCK_N + uint32_t(c0), r);
#pragma unroll
          for (int v = 0; v < 4; ++v) mk[v] = mk_next[v];
          if (col0 + CHUNK < N) load_mask(mk_next, col0 + CHUNK);
          tmem_ld_wait();
        }
        if (c0 + CHUNK >= BLOCK_N || col0 + CHUNK >= N) {
          // last TMEM read of this tile is complete: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
        if (!live) break;
        epilogue_chunk<CHUNK>(ep, r, mk, have_mask, mrow, have_bias, s_bias, c0, row, row_ok, col0, N, lane, 0, s_sum, s_sq);
        if (col0 + CHUNK >= N) break;
      }
    }
    if (stat_mode) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      flush_stats();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// ----------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 matrix with `rows` rows of `cols` contiguous elements (pitch ld); box = box_rows x 64 cols,
// 128B swizzle, out-of-bounds elements read as 0.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return -1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : int(r);
}

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                     const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// x [N, H, W, C] bf16 (dense NHWC) as an im2col tensor map: boxes of 128 output pixels x 64 channels, 128B swizzle, zero fill.
// Bounding box: lower corner = -pad, upper corner = pad - (filter - 1) * dilation, so that the box walk visits exactly the
// P x Q receptive-field corners of every image; the filter tap is added per load (instruction offsets).
static int make_tmap_im2col_bf16(CUtensorMap* map, const void* x, int N, int H, int W, int C, int R, int S, int stride, int pad, int dil) {
  static PFN_encodeIm2col enc = nullptr;
  if (enc == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) != cudaSuccess || p == nullptr) return -1;
    enc = reinterpret_cast<PFN_encodeIm2col>(p);
  }
  cuuint64_t gdim[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
  cuuint64_t gstride[3] = {cuuint64_t(C) * 2, cuuint64_t(W) * C * 2, cuuint64_t(H) * W * C * 2};
  int lower[2] = {-pad, -pad};                                          // (w, h)
  int upper[2] = {pad - (S - 1) * dil, pad - (R - 1) * dil};
  cuuint32_t estr[4] = {1, cuuint32_t(stride), cuuint32_t(stride), 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstride, lower, upper, BLOCK_K, BLOCK_M, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return int(r);
  // (driver quirk handled the way CUTLASS does - cute/atom/copy_traits_sm90_im2col.hpp: descriptors of tensors below 128 KB get
  //  a bit set by drivers <= 13.1 that the im2col unit then trips over)
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && int64_t(N) * H * W * C * 2 < 131072)
    reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return 0;
}

// Tile width: the smallest power of two covering N, then shrunk further until ~64 CTAs are in
// flight (tiny problems are bound by how much ONE SM can pull from L2, so they want many small
// tiles); big problems keep 128-wide tiles.
int gemm_pick_block_n(int M, int N, int b_mn) {
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  const int min_bn = b_mn ? 64 : 16;
  int bn = 128;
  while (bn > min_bn && bn / 2 >= N) bn >>= 1;
  while (bn > min_bn && tiles_m * ((N + bn - 1) / bn) < 64) bn >>= 1;
  return bn;
}

template <int BN, bool A_MN, bool B_MN>
static cudaError_t launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, const GemmWindow& win,
                             const GemmPeerB& peer, int M, int N, int K, int split_k, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
  int kb_per = (total_kb + split_k - 1) / split_k;
  int splits = (total_kb + kb_per - 1) / kb_per;      // no empty splits
  dim3 grid((N + BN - 1) / BN, (M + BLOCK_M - 1) / BLOCK_M, splits);
  return launch_kernel(gemm_bf16_tn_kernel<BN, A_MN, B_MN>, grid, dim3(GEMM_THREADS), S::TOTAL, stream, ta, tb, ep, win, peer, M, N, K, kb_per);
}

template <int BN, bool A_MN, bool B_MN>
static cudaError_t launch_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                                     cudaStream_t stream) {
  using S = GemmSmem<BN>;
  static bool attr_set = false;
  static int num_sms = 148;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_persistent_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_kernel(gemm_bf16_persistent_kernel<BN, A_MN, B_MN>, dim3(grid), dim3(GEMM_THREADS), S::TOTAL, stream, ta, tb, ep, M, N, K);
}

template <bool A_MN, bool B_MN>
static cudaError_t dispatch_persistent(int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M,
                                       int N, int K, cudaStream_t stream) {
  switch (block_n) {
    case 64: return launch_persistent<64, A_MN, B_MN>(ta, tb, ep, M, N, K, stream);
    case 128: return launch_persistent<128, A_MN, B_MN>(ta, tb, ep, M, N, K, stream);
    case 256: return launch_persistent<256, A_MN, B_MN>(ta, tb, ep, M, N, K, stream);
  }
  return cudaErrorInvalidValue;
}

template <bool A_MN, bool B_MN>
static cudaError_t dispatch_bn(int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep,
                               const GemmWindow& win, const GemmPeerB& peer, int M, int N, int K, int split_k,
                               cudaStream_t stream) {
  switch (block_n) {
    case 16: if constexpr (!B_MN) return launch_bn<16, A_MN, B_MN>(ta, tb, ep, win, peer, M, N, K, split_k, stream); break;
    case 32: if constexpr (!B_MN) return launch_bn<32, A_MN, B_MN>(ta, tb, ep, win, peer, M, N, K, split_k, stream); break;
    case 64: return launch_bn<64, A_MN, B_MN>(ta, tb, ep, win, peer, M, N, K, split_k, stream);
    case 128: return launch_bn<128, A_MN, B_MN>(ta, tb, ep, win, peer, M, N, K, split_k, stream);
    case 256: return launch_bn<256, A_MN, B_MN>(ta, tb, ep, win, peer, M, N, K, split_k, stream);
  }
  return cudaErrorInvalidValue;
}

// D[M,N] = A * B^T.  a_mn=0: A stored [M,K] (pitch lda); a_mn=1: A stored [K,M].  Same for B with N.
// bf16, base pointers 16-byte aligned, pitches multiples of 8 elements.
int gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int M, int N, int K,
              const GemmEpilogue* ep, const GemmWindow* winp, int64_t a_rows_total, int64_t b_rows_total, int block_n,
              int split_k, cudaStream_t stream, const GemmPeerB* peerp) {
  GemmWindow win{};
  if (winp != nullptr) win = *winp;
  GemmPeerB peer{};
  if (peerp != nullptr) peer = *peerp;
  if (peer.world > 0) {
    if (b_mn) return -6;                       // the gathered operand is K-major
    if (block_n <= 0) block_n = N <= 16 ? 16 : 32;
    if (block_n > 32) return -7;
  }
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda & 7) || (reinterpret_cast<uintptr_t>(A) & 15)) return -2;
  if ((peerp == nullptr || peerp->world == 0) && ((ldb & 7) || (reinterpret_cast<uintptr_t>(B) & 15))) return -2;
  if (block_n <= 0) block_n = gemm_pick_block_n(M, N, b_mn);
  if (ep->loss_kind != 0 && N > (block_n >= 32 ? 32 : 16)) return -5;   // fused loss needs the row in one chunk
  if (b_mn && block_n < 64) block_n = 64;
  if (split_k <= 0) split_k = 1;
  if (split_k > 1 && !ep->atomic_f32) return -3;
  // Big plain products (forward of wide Linear / im2col'ed convolutions): CTA pairs with tcgen05.mma.cta_group::2 (gemm2cta_sm100.cu)
  if (g_pair_mode != 0 && !a_mn && !b_mn && peer.world == 0 && win.flags == 0 && split_k == 1 && ep->out_bf16 != nullptr &&
      ep->mask_src == nullptr && ep->out_t_bf16 == nullptr && ep->out_f32 == nullptr && ep->out_t_f32 == nullptr && ep->colsum == nullptr &&
      ep->colsumsq == nullptr && ep->loss_kind == 0 && ep->push_world == 0 && ep->prof == nullptr && M >= 512 && N >= 256 && K >= 256 &&
      // (enough 256 x 256 tiles to keep most of the 74 pairs busy: below that the single-CTA kernels' smaller tiles spread further)
      int64_t((M + 255) / 256) * ((N + 255) / 256) >= 48 && (ep->ld_out & 7) == 0) {
    const int rc2 = gemm_bf16_2cta(A, lda, B, ldb, ep->out_bf16, ep->ld_out, M, N, K, ep->bias, ep->alpha, ep->relu, stream);
    if (rc2 != -2) return rc2;
  }
  CUtensorMap ta, tb;
  // windowed operands are described over the whole resident matrix (rows_total), the kernel adds the offset
  const uint64_t a_rows = (win.flags & 1) && a_rows_total > 0 ? uint64_t(a_rows_total) : uint64_t(M);
  const uint64_t b_krows = (win.flags & 2) && b_rows_total > 0 ? uint64_t(b_rows_total) : uint64_t(K);
  int rc = a_mn ? make_tmap_bf16_2d(&ta, A, K, M, lda, BLOCK_K) : make_tmap_bf16_2d(&ta, A, a_rows, K, lda, BLOCK_M);
  if (rc) return rc;
  if (peer.world > 0) tb = ta;                 // unused by the kernel in K-AG mode
  else rc = b_mn ? make_tmap_bf16_2d(&tb, B, b_krows, N, ldb, BLOCK_K) : make_tmap_bf16_2d(&tb, B, N, K, ldb, block_n);
  if (rc) return rc;
  cudaError_t e;
  // Large problems: persistent CTAs with a double-buffered TMEM accumulator (epilogue overlaps the next tile's MMAs).
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + block_n - 1) / block_n);
  const bool persistent = g_persistent_mode != 0 && block_n >= 64 && split_k == 1 && peer.world == 0 && win.flags == 0 &&
                          ep->loss_kind == 0 && ep->prof == nullptr && (g_persistent_mode == 2 || tiles > 148);
  if (persistent) {
    if (a_mn && b_mn) e = dispatch_persistent<true, true>(block_n, ta, tb, *ep, M, N, K, stream);
    else if (a_mn) e = dispatch_persistent<true, false>(block_n, ta, tb, *ep, M, N, K, stream);
    else if (b_mn) e = dispatch_persistent<false, true>(block_n, ta, tb, *ep, M, N, K, stream);
    else e = dispatch_persistent<false, false>(block_n, ta, tb, *ep, M, N, K, stream);
    return e == cudaSuccess ? 0 : 1000 + int(e);
  }
  if (a_mn && b_mn) e = dispatch_bn<true, true>(block_n, ta, tb, *ep, win, peer, M, N, K, split_k, stream);
  else if (a_mn) e = dispatch_bn<true, false>(block_n, ta, tb, *ep, win, peer, M, N, K, split_k, stream);
  else if (b_mn) e = dispatch_bn<false, true>(block_n, ta, tb, *ep, win, peer, M, N, K, split_k, stream);
  else e = dispatch_bn<false, false>(block_n, ta, tb, *ep, win, peer, M, N, K, split_k, stream);
  return e == cudaSuccess ? 0 : 1000 + int(e);
}

int conv2d_fwd_implicit(const void* x, int N, int H, int W, int C, const void* w, int64_t ldw, int Cout, int R, int S, int stride, int pad,
                        int dil, const GemmEpilogue* ep, int block_n, cudaStream_t stream) {
  if (C % BLOCK_K != 0 || stride < 1 || stride > 8 || dil < 1 || R < 1 || S < 1) return -2;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (ldw & 7)) return -2;
  const int P = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1, Q = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  if (P < 1 || Q < 1 || pad > 127 || (S - 1) * dil - pad > 128 || (R - 1) * dil - pad > 128) return -2;
  const int64_t M64 = int64_t(N) * P * Q;
  if (M64 > (int64_t(1) << 30)) return -2;
  const int M = int(M64), K = R * S * C;
  if (ep->loss_kind != 0) return -5;
  CUtensorMap ta, tb;
  int rc = make_tmap_im2col_bf16(&ta, x, N, H, W, C, R, S, stride, pad, dil);
  if (rc) return rc;
  if (block_n <= 0) block_n = gemm_pick_block_n(M, Cout, 0);
  rc = make_tmap_bf16_2d(&tb, w, uint64_t(Cout), uint64_t(K), uint64_t(ldw), uint32_t(block_n));
  if (rc) return rc;
  GemmWindow win{};
  win.flags = 8;
  win.conv_cb = C / BLOCK_K; win.conv_S = S; win.conv_P = P; win.conv_Q = Q; win.conv_stride = stride; win.conv_pad = pad; win.conv_dil = dil;
  GemmPeerB peer{};
  cudaError_t e = dispatch_bn<false, false>(block_n, ta, tb, *ep, win, peer, M, Cout, K, 1, stream);
  return e == cudaSuccess ? 0 : 1000 + int(e);
}

}  // namespace st
