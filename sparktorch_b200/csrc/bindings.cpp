// Python bindings (torch extension) for the sm_100a kernels + the native step
// runner (CUDA-graph capture/replay loop with pinned-memory input streaming).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <optional>
#include <vector>

#include "gemm_sm100.h"
#include "kernels.h"
#include "mlp_step.h"

namespace {

using torch::Tensor;
using OptT = std::optional<Tensor>;

cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, " failed with code ", rc, " (", rc >= 1000 ? cudaGetErrorString(cudaError_t(rc - 1000)) : "", ")");
}

const __nv_bfloat16* bf16p(const Tensor& t) { return reinterpret_cast<const __nv_bfloat16*>(t.data_ptr()); }
__nv_bfloat16* bf16pm(const Tensor& t) { return reinterpret_cast<__nv_bfloat16*>(t.data_ptr()); }

void req_2d(const Tensor& t, c10::ScalarType dt, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.dim() == 2 && t.stride(1) == 1 && t.scalar_type() == dt, name, ": need a CUDA 2-D tensor with unit inner stride of the right dtype");
}

// D[M,N] = alpha * A * B^T (+ fused epilogue).  a_mn: A is stored [K, M]; b_mn: B is stored [K, N].
void gemm_bf16(const Tensor& A, bool a_mn, const Tensor& B, bool b_mn, int64_t M, int64_t N, int64_t K, double alpha,
               const OptT& bias, bool relu, const OptT& mask_src, const OptT& out_bf16, const OptT& out_t_bf16,
               const OptT& out_f32, const OptT& out_t_f32, bool atomic_f32, const OptT& colsum, int64_t block_n,
               int64_t split_k, int64_t loss_kind, const OptT& labels, const OptT& target, const OptT& loss_slot,
               double loss_scale, const OptT& win_state, int64_t window_rows, int64_t n_windows, int64_t win_flags,
               const OptT& prof, const OptT& peer_table, int64_t peer_shard_elems, int64_t peer_w_off,
               int64_t peer_bias_off, bool peer_fill_local, const OptT& push_table, int64_t push_me,
               int64_t push_mask, const OptT& push_state, const OptT& colsumsq) {
  req_2d(A, at::kBFloat16, "A");
  req_2d(B, at::kBFloat16, "B");
  c10::cuda::CUDAGuard guard(A.device());
  st::GemmEpilogue ep;
  ep.alpha = float(alpha);
  ep.relu = relu ? 1 : 0;
  ep.atomic_f32 = atomic_f32 ? 1 : 0;
  if (bias) { TORCH_CHECK(bias->scalar_type() == at::kFloat && bias->numel() >= N); ep.bias = bias->data_ptr<float>(); }
  if (mask_src) {
    req_2d(*mask_src, at::kBFloat16, "mask_src");
    TORCH_CHECK(mask_src->stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(mask_src->data_ptr()) & 15) == 0, "mask_src: pitch must be a multiple of 8 elements, base 16-byte aligned");
    ep.mask_src = bf16p(*mask_src); ep.ld_mask = mask_src->stride(0);
  }
  if (out_bf16) {
    req_2d(*out_bf16, at::kBFloat16, "out_bf16");
    TORCH_CHECK(out_bf16->stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(out_bf16->data_ptr()) & 15) == 0, "out_bf16: pitch must be a multiple of 8 elements, base 16-byte aligned");
    ep.out_bf16 = bf16pm(*out_bf16); ep.ld_out = out_bf16->stride(0);
  }
  if (out_t_bf16) { req_2d(*out_t_bf16, at::kBFloat16, "out_t_bf16"); ep.out_t_bf16 = bf16pm(*out_t_bf16); ep.ld_out_t = out_t_bf16->stride(0); }
  if (out_f32) { req_2d(*out_f32, at::kFloat, "out_f32"); ep.out_f32 = out_f32->data_ptr<float>(); ep.ld_f32 = out_f32->stride(0); }
  if (out_t_f32) { req_2d(*out_t_f32, at::kFloat, "out_t_f32"); ep.out_t_f32 = out_t_f32->data_ptr<float>(); ep.ld_f32_t = out_t_f32->stride(0); }
  if (colsum) { TORCH_CHECK(colsum->scalar_type() == at::kFloat && colsum->numel() >= N); ep.colsum = colsum->data_ptr<float>(); }
  if (colsumsq) { TORCH_CHECK(colsumsq->scalar_type() == at::kFloat && colsumsq->numel() >= N && !bias.has_value()); ep.colsumsq = colsumsq->data_ptr<float>(); }
  ep.loss_kind = int(loss_kind);
  if (loss_kind != 0) {
    TORCH_CHECK(loss_slot.has_value() && loss_slot->scalar_type() == at::kFloat, "fused loss needs a float loss slot");
    ep.loss_slot = loss_slot->data_ptr<float>();
    ep.loss_scale = float(loss_scale);
    if (loss_kind == 1) { TORCH_CHECK(labels.has_value() && labels->scalar_type() == at::kInt); ep.ce_labels = labels->data_ptr<int32_t>(); }
    if (loss_kind == 2) { TORCH_CHECK(target.has_value()); req_2d(*target, at::kFloat, "target"); ep.mse_target = target->data_ptr<float>(); ep.ld_target = target->stride(0); }
  }
  if (prof) { TORCH_CHECK(prof->scalar_type() == at::kLong); ep.prof = reinterpret_cast<long long*>(prof->data_ptr()); }
  st::GemmWindow win;
  int64_t a_rows_total = 0, b_rows_total = 0;
  if (win_state && win_flags != 0) {
    win.state = reinterpret_cast<const st::StepState*>(win_state->data_ptr());
    win.window_rows = window_rows; win.n_windows = n_windows > 0 ? n_windows : 1; win.flags = int(win_flags);
    if (win_flags & 1) a_rows_total = A.size(0);
    if (win_flags & 2) b_rows_total = B.size(0);
  }
  if (push_table && push_mask != 0) {
    TORCH_CHECK(push_table->is_cuda() && push_table->scalar_type() == at::kLong && push_table->numel() % 2 == 0 &&
                push_table->numel() <= 16 && push_state.has_value(), "push: device int64 table [2][world] + state");
    ep.push_world = int(push_table->numel() / 2); ep.push_me = int(push_me); ep.push_mask = int(push_mask);
    ep.push_delta = reinterpret_cast<const long long*>(push_table->data_ptr());
    ep.push_state = reinterpret_cast<const st::StepState*>(push_state->data_ptr());
  }
  st::GemmPeerB peer;
  if (peer_table) {
    TORCH_CHECK(peer_table->is_cuda() && peer_table->scalar_type() == at::kLong && peer_table->numel() <= 8 && peer_shard_elems > 0,
                "K-AG: device int64 table of up to 8 owner pointers");
    peer.world = int(peer_table->numel());
    peer.param = reinterpret_cast<const float* const*>(peer_table->data_ptr());
    peer.shard_elems = peer_shard_elems; peer.w_off = peer_w_off; peer.bias_off = peer_bias_off;
    if (peer_fill_local) { peer.wb_local = bf16pm(B); peer.ld_wb = B.stride(0); }
  }
  check(st::gemm_bf16(A.data_ptr(), A.stride(0), a_mn ? 1 : 0, B.data_ptr(), B.stride(0), b_mn ? 1 : 0, int(M), int(N),
                      int(K), &ep, &win, a_rows_total, b_rows_total, int(block_n), int(split_k), cur_stream(), &peer),
        "gemm_bf16");
}

static const st::StepState* win_ptr(const OptT& win_state) {
  return win_state ? reinterpret_cast<const st::StepState*>(win_state->data_ptr()) : nullptr;
}

// win_state/window_rows/n_windows: labels (targets) are the resident partition's, the batch is the current window
void softmax_ce(const Tensor& logits, const Tensor& labels, int64_t C, const Tensor& loss_slot, const OptT& dz,
                const OptT& dzt, const OptT& db, int64_t pad_cols, const OptT& win_state, int64_t window_rows, int64_t n_windows) {
  req_2d(logits, at::kFloat, "logits");
  TORCH_CHECK(labels.scalar_type() == at::kInt && loss_slot.scalar_type() == at::kFloat);
  TORCH_CHECK(labels.numel() >= (win_state ? std::max<int64_t>(1, n_windows) * window_rows : logits.size(0)), "labels too short");
  c10::cuda::CUDAGuard guard(logits.device());
  check(st::softmax_ce(logits.data_ptr<float>(), logits.stride(0), labels.data_ptr<int32_t>(), int(logits.size(0)), int(C),
                       loss_slot.data_ptr<float>(), dz ? bf16pm(*dz) : nullptr, dz ? dz->stride(0) : 0,
                       dzt ? bf16pm(*dzt) : nullptr, dzt ? dzt->stride(0) : 0, db ? db->data_ptr<float>() : nullptr,
                       int(pad_cols), cur_stream(), win_ptr(win_state), window_rows, n_windows),
        "softmax_ce");
}

void mse_loss(const Tensor& pred, const Tensor& target, int64_t D, const Tensor& loss_slot, const OptT& dz, const OptT& dzt,
              const OptT& db, int64_t pad_cols, const OptT& win_state, int64_t window_rows, int64_t n_windows) {
  req_2d(pred, at::kFloat, "pred");
  req_2d(target, at::kFloat, "target");
  c10::cuda::CUDAGuard guard(pred.device());
  check(st::mse_loss(pred.data_ptr<float>(), pred.stride(0), target.data_ptr<float>(), target.stride(0), int(pred.size(0)),
                     int(D), loss_slot.data_ptr<float>(), dz ? bf16pm(*dz) : nullptr, dz ? dz->stride(0) : 0,
                     dzt ? bf16pm(*dzt) : nullptr, dzt ? dzt->stride(0) : 0, db ? db->data_ptr<float>() : nullptr, int(pad_cols),
                     cur_stream(), win_ptr(win_state), window_rows, n_windows),
        "mse_loss");
}

void gather_batch(const Tensor& src, int64_t n_rows, const OptT& tgt_src, int64_t tgt_dim, int64_t mode, int64_t ring,
                  int64_t seed, int64_t B, int64_t D, const Tensor& x, const OptT& xt, const OptT& tgt_out, const OptT& state) {
  req_2d(src, at::kBFloat16, "src");
  req_2d(x, at::kBFloat16, "x");
  c10::cuda::CUDAGuard guard(src.device());
  st::GatherArgs a{};
  a.src = bf16p(src); a.ld_src = src.stride(0); a.n_rows = n_rows;
  a.tgt_src = tgt_src ? tgt_src->data_ptr() : nullptr; a.tgt_dim = int(tgt_dim);
  a.mode = int(mode); a.ring = int(ring > 0 ? ring : 1); a.seed = uint64_t(seed);
  a.B = int(B); a.D = int(D);
  a.x = bf16pm(x); a.ldx = x.stride(0);
  if (xt) { req_2d(*xt, at::kBFloat16, "xt"); a.xt = bf16pm(*xt); a.ldxt = xt->stride(0); }
  a.tgt_out = tgt_out ? tgt_out->data_ptr() : nullptr;
  a.state = state ? reinterpret_cast<const st::StepState*>(state->data_ptr()) : nullptr;
  check(st::gather_batch(a, cur_stream()), "gather_batch");
}

void transpose_bf16(const Tensor& src, int64_t rows, int64_t cols, const Tensor& dst) {
  req_2d(src, at::kBFloat16, "src");
  req_2d(dst, at::kBFloat16, "dst");
  c10::cuda::CUDAGuard guard(src.device());
  check(st::transpose_bf16(bf16p(src), src.stride(0), int(rows), int(cols), bf16pm(dst), dst.stride(0), cur_stream()), "transpose_bf16");
}

// ------------------------------------------------------------------ convolution helpers
void im2col_nhwc(const Tensor& x, int64_t B, int64_t H, int64_t W, int64_t Cc, int64_t sb, int64_t sh, int64_t sw, int64_t sc,
                 int64_t kh, int64_t kw, int64_t stride, int64_t pad, const Tensor& col, const OptT& state, int64_t window_rows,
                 int64_t n_windows) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16);
  req_2d(col, at::kBFloat16, "col");
  c10::cuda::CUDAGuard guard(x.device());
  st::Im2colArgs a{};
  a.x = bf16p(x); a.sb = sb; a.sh = sh; a.sw = sw; a.sc = sc;
  a.state = state ? reinterpret_cast<const st::StepState*>(state->data_ptr()) : nullptr;
  a.window_rows = window_rows; a.n_windows = n_windows > 0 ? n_windows : 1;
  a.B = int(B); a.H = int(H); a.W = int(W); a.C = int(Cc);
  a.kh = int(kh); a.kw = int(kw); a.stride = int(stride); a.pad = int(pad);
  a.OH = int((H + 2 * pad - kh) / stride + 1); a.OW = int((W + 2 * pad - kw) / stride + 1);
  a.K = int(Cc * kh * kw); a.col = bf16pm(col); a.ld_col = col.stride(0);
  TORCH_CHECK(col.size(0) >= B * a.OH * a.OW && col.stride(0) >= a.K, "col buffer too small");
  check(st::im2col_nhwc(a, cur_stream()), "im2col_nhwc");
}

void flip_conv_weight(const Tensor& w, int64_t Cout, int64_t Cin, int64_t kh, int64_t kw, const Tensor& out) {
  req_2d(w, at::kBFloat16, "w");
  req_2d(out, at::kBFloat16, "out");
  c10::cuda::CUDAGuard guard(w.device());
  check(st::flip_conv_weight(bf16p(w), w.stride(0), int(Cout), int(Cin), int(kh), int(kw), bf16pm(out), out.stride(0), cur_stream()),
        "flip_conv_weight");
}

void maxpool_dropout(bool backward, const Tensor& x_or_dx, int64_t ld_x, const Tensor& y_or_dy, const Tensor& idx, int64_t B,
                     int64_t H, int64_t W, int64_t Cc, int64_t k, double p_drop, bool train, int64_t seed, const OptT& state,
                     bool y_nhwc, const OptT& relu_x, const OptT& db) {
  req_2d(y_or_dy, at::kBFloat16, "y");
  TORCH_CHECK(x_or_dx.scalar_type() == at::kBFloat16 && idx.scalar_type() == at::kByte);
  c10::cuda::CUDAGuard guard(x_or_dx.device());
  st::PoolArgs a{};
  a.B = int(B); a.H = int(H); a.W = int(W); a.C = int(Cc); a.k = int(k); a.PH = int(H / k); a.PW = int(W / k);
  a.p_drop = float(p_drop); a.train = train; a.seed = uint64_t(seed);
  a.state = state ? reinterpret_cast<const st::StepState*>(state->data_ptr()) : nullptr;
  a.idx = idx.data_ptr<uint8_t>();
  a.y_nhwc = y_nhwc;
  a.relu_x = relu_x ? bf16p(*relu_x) : nullptr;
  a.db = db ? db->data_ptr<float>() : nullptr;
  if (!backward) {
    a.x = bf16p(x_or_dx); a.ld_x = ld_x; a.y = bf16pm(y_or_dy); a.ld_y = y_or_dy.stride(0);
    check(st::maxpool_dropout_fwd(a, cur_stream()), "maxpool_dropout_fwd");
  } else {
    a.dx = bf16pm(x_or_dx); a.ld_dx = ld_x; a.dy = bf16p(y_or_dy); a.ld_dy = y_or_dy.stride(0);
    check(st::maxpool_dropout_bwd(a, cur_stream()), "maxpool_dropout_bwd");
  }
}

// ------------------------------------------------------------------ residual conv nets (norm.cu)
static const __nv_bfloat16* bfo(const OptT& t) { return t ? bf16p(*t) : nullptr; }
static int64_t ldo(const OptT& t) { return t ? t->stride(0) : 0; }

void im2col_cl(const Tensor& x, int64_t B, int64_t H, int64_t W, int64_t Cc, int64_t kh, int64_t kw, int64_t stride, int64_t pad,
               int64_t dil, int64_t OH, int64_t OW, const Tensor& col) {
  req_2d(x, at::kBFloat16, "x");
  req_2d(col, at::kBFloat16, "col");
  c10::cuda::CUDAGuard guard(x.device());
  st::Im2colClArgs a{};
  a.x = bf16p(x); a.ld_x = x.stride(0);
  a.B = int(B); a.H = int(H); a.W = int(W); a.C = int(Cc); a.kh = int(kh); a.kw = int(kw);
  a.stride = int(stride); a.pad = int(pad); a.dil = int(dil < 1 ? 1 : dil); a.OH = int(OH); a.OW = int(OW);
  a.col = bf16pm(col); a.ld_col = col.stride(0);
  TORCH_CHECK(x.size(0) >= B * H * W && x.size(1) >= Cc, "im2col_cl: x too small");
  TORCH_CHECK(col.size(0) >= B * OH * OW && col.size(1) >= kh * kw * Cc, "im2col_cl: col too small");
  check(st::im2col_cl(a, cur_stream()), "im2col_cl");
}

void flip_conv_weight_cl(const Tensor& w, int64_t Cout, int64_t Cin, int64_t kh, int64_t kw, const Tensor& out) {
  req_2d(w, at::kBFloat16, "w");
  req_2d(out, at::kBFloat16, "out");
  TORCH_CHECK(w.size(0) >= Cout && w.size(1) >= kh * kw * Cin && out.size(0) >= Cin && out.size(1) >= kh * kw * Cout);
  c10::cuda::CUDAGuard guard(w.device());
  check(st::flip_conv_weight_cl(bf16p(w), w.stride(0), int(Cout), int(Cin), int(kh), int(kw), bf16pm(out), out.stride(0), cur_stream()),
        "flip_conv_weight_cl");
}

void bn_stats(const Tensor& y, int64_t M, int64_t Cc, const Tensor& sums) {
  req_2d(y, at::kBFloat16, "y");
  TORCH_CHECK(sums.scalar_type() == at::kFloat && sums.numel() >= 2 * Cc && y.size(0) >= M && y.size(1) >= Cc);
  c10::cuda::CUDAGuard guard(y.device());
  check(st::bn_stats(bf16p(y), y.stride(0), M, int(Cc), sums.data_ptr<float>(), cur_stream()), "bn_stats");
}

void bn_finalize(const Tensor& sums, int64_t M, int64_t Cc, double eps, double momentum, const OptT& running_mean,
                 const OptT& running_var, const Tensor& stats, bool from_running, bool update_running) {
  TORCH_CHECK(sums.scalar_type() == at::kFloat && stats.scalar_type() == at::kFloat && stats.numel() >= 2 * Cc);
  if (running_mean) TORCH_CHECK(running_mean->scalar_type() == at::kFloat && running_var->scalar_type() == at::kFloat &&
                                running_mean->is_contiguous() && running_var->is_contiguous() && running_mean->numel() >= Cc);
  TORCH_CHECK(!from_running || running_mean.has_value());
  c10::cuda::CUDAGuard guard(sums.device());
  check(st::bn_finalize(sums.data_ptr<float>(), M, int(Cc), float(eps), float(momentum),
                        running_mean ? running_mean->data_ptr<float>() : nullptr, running_var ? running_var->data_ptr<float>() : nullptr,
                        stats.data_ptr<float>(), from_running, update_running, cur_stream()), "bn_finalize");
}

static st::BnArgs bn_args(int64_t M, int64_t Cc, bool relu, const Tensor& y, const Tensor& z, const Tensor& stats, const Tensor& gamma) {
  req_2d(y, at::kBFloat16, "y");
  req_2d(z, at::kBFloat16, "z");
  TORCH_CHECK(y.size(0) >= M && y.size(1) >= Cc && z.size(0) >= M && z.size(1) >= Cc && stats.numel() >= 2 * Cc && gamma.numel() >= Cc);
  TORCH_CHECK(stats.scalar_type() == at::kFloat && gamma.scalar_type() == at::kFloat);
  st::BnArgs a{};
  a.M = M; a.C = int(Cc); a.relu = relu;
  a.y = bf16p(y); a.ld = y.stride(0);
  a.z = bf16pm(z); a.ld_z = z.stride(0);
  a.stats = stats.data_ptr<float>();
  a.gamma = gamma.data_ptr<float>();
  return a;
}

void bn_apply(const Tensor& y, int64_t M, int64_t Cc, const Tensor& stats, const Tensor& gamma, const Tensor& beta, const Tensor& z,
              bool relu, const OptT& residual, const OptT& sums, double eps, double momentum, const OptT& running_mean,
              const OptT& running_var) {
  c10::cuda::CUDAGuard guard(y.device());
  st::BnArgs a = bn_args(M, Cc, relu, y, z, stats, gamma);
  if (sums) {            // statistics come straight from the GEMM epilogue's column sums: finalize inline
    TORCH_CHECK(sums->scalar_type() == at::kFloat && sums->numel() >= 2 * Cc);
    a.sums = sums->data_ptr<float>();
    a.stats_out = stats.data_ptr<float>();
    a.eps = float(eps); a.momentum = float(momentum);
    if (running_mean) {
      TORCH_CHECK(running_var.has_value() && running_mean->scalar_type() == at::kFloat && running_var->scalar_type() == at::kFloat &&
                  running_mean->is_contiguous() && running_var->is_contiguous() && running_mean->numel() >= Cc);
      a.running_mean = running_mean->data_ptr<float>(); a.running_var = running_var->data_ptr<float>();
    }
  }
  TORCH_CHECK(beta.scalar_type() == at::kFloat && beta.numel() >= Cc);
  a.beta = beta.data_ptr<float>();
  if (residual) { req_2d(*residual, at::kBFloat16, "residual"); TORCH_CHECK(residual->size(0) >= M && residual->size(1) >= Cc); }
  a.residual = bfo(residual); a.ld_res = ldo(residual);
  check(st::bn_apply(a, cur_stream()), "bn_apply");
}

// reduce=true: dgamma/dbeta += ...;  reduce=false: dy (and g_out) from the finished dgamma/dbeta
void bn_backward(bool reduce, const Tensor& y, int64_t M, int64_t Cc, const Tensor& stats, const Tensor& gamma, const Tensor& z,
                 bool relu, const Tensor& dz_a, const OptT& dz_b, const Tensor& dgamma, const Tensor& dbeta, const OptT& dy,
                 const OptT& g_out) {
  c10::cuda::CUDAGuard guard(y.device());
  st::BnArgs a = bn_args(M, Cc, relu, y, z, stats, gamma);
  req_2d(dz_a, at::kBFloat16, "dz_a");
  TORCH_CHECK(dz_a.size(0) >= M && dz_a.size(1) >= Cc);
  a.dz_a = bf16p(dz_a); a.ld_dz = dz_a.stride(0);
  if (dz_b) { req_2d(*dz_b, at::kBFloat16, "dz_b"); TORCH_CHECK(dz_b->stride(0) == dz_a.stride(0) && dz_b->size(0) >= M, "dz_b pitch"); a.dz_b = bf16p(*dz_b); }
  TORCH_CHECK(dgamma.scalar_type() == at::kFloat && dbeta.scalar_type() == at::kFloat && dgamma.numel() >= Cc && dbeta.numel() >= Cc);
  a.dgamma = dgamma.data_ptr<float>(); a.dbeta = dbeta.data_ptr<float>();
  if (reduce) {
    check(st::bn_bwd_reduce(a, cur_stream()), "bn_bwd_reduce");
  } else {
    TORCH_CHECK(dy.has_value());
    req_2d(*dy, at::kBFloat16, "dy");
    TORCH_CHECK(dy->size(0) >= M && dy->size(1) >= Cc);
    a.dy = bf16pm(*dy); a.ld_dy = dy->stride(0);
    if (g_out) { req_2d(*g_out, at::kBFloat16, "g_out"); TORCH_CHECK(g_out->stride(0) == dz_a.stride(0) && g_out->size(0) >= M, "g_out pitch"); a.g_out = bf16pm(*g_out); }
    check(st::bn_bwd_apply(a, cur_stream()), "bn_bwd_apply");
  }
}

void maxpool_cl(bool backward, const Tensor& x_or_dx, const Tensor& y_or_dy, const Tensor& idx, int64_t B, int64_t H, int64_t W,
                int64_t Cc, int64_t k, int64_t s, int64_t p, const OptT& dy_b) {
  req_2d(x_or_dx, at::kBFloat16, "x");
  req_2d(y_or_dy, at::kBFloat16, "y");
  TORCH_CHECK(idx.scalar_type() == at::kByte);
  c10::cuda::CUDAGuard guard(x_or_dx.device());
  st::Pool2Args a{};
  a.B = int(B); a.H = int(H); a.W = int(W); a.C = int(Cc); a.k = int(k); a.s = int(s); a.p = int(p);
  a.PH = int((H + 2 * p - k) / s + 1); a.PW = int((W + 2 * p - k) / s + 1);
  TORCH_CHECK(x_or_dx.size(0) >= B * H * W && y_or_dy.size(0) >= B * a.PH * a.PW && idx.numel() >= B * a.PH * a.PW * Cc);
  a.ld_x = x_or_dx.stride(0); a.ld_y = y_or_dy.stride(0);
  a.idx = idx.data_ptr<uint8_t>();
  if (!backward) { a.x = bf16p(x_or_dx); a.y = bf16pm(y_or_dy); }
  else {
    a.dx = bf16pm(x_or_dx); a.dy = bf16p(y_or_dy);
    if (dy_b) { req_2d(*dy_b, at::kBFloat16, "dy_b"); TORCH_CHECK(dy_b->stride(0) == y_or_dy.stride(0) && dy_b->size(0) >= B * a.PH * a.PW); a.dy_b = bf16p(*dy_b); }
  }
  check(st::maxpool_cl(a, backward, cur_stream()), "maxpool_cl");
}

void avgpool_cl(bool backward, const Tensor& x_or_dx, const Tensor& y_or_dy, int64_t B, int64_t HW, int64_t Cc) {
  req_2d(x_or_dx, at::kBFloat16, "x");
  req_2d(y_or_dy, at::kBFloat16, "y");
  TORCH_CHECK(x_or_dx.size(0) >= B * HW && y_or_dy.size(0) >= B && x_or_dx.size(1) >= Cc && y_or_dy.size(1) >= Cc);
  c10::cuda::CUDAGuard guard(x_or_dx.device());
  check(st::avgpool_cl(bf16pm(x_or_dx), x_or_dx.stride(0), bf16pm(y_or_dy), y_or_dy.stride(0), int(B), int(HW), int(Cc), backward,
                       cur_stream()), "avgpool_cl");
}

// ------------------------------------------------------------------ fused all-reduce + optimizer
class Updater {
 public:
  Updater(int64_t world, int64_t rank, std::vector<int64_t> grad_ptrs, std::vector<int64_t> param_ptrs,
          std::vector<int64_t> wbf16_ptrs, std::vector<int64_t> flag_ptrs, int64_t mc_grad_ptr, const OptT& m, const OptT& v,
          const Tensor& state, const OptT& loss_hist, int64_t host_mailbox_ptr, const Tensor& segs, int64_t numel,
          int64_t tail, int64_t kind, double lr, double beta1, double beta2, double eps, double weight_decay, double momentum,
          double dampening, bool nesterov, bool maximize, int64_t patience, bool zero_grads, int64_t recv_local_ptr,
          int64_t recv_stride, int64_t mc_param_ptr, int64_t mc_wbf16_ptr)
      : state_(state), segs_(segs), m_(m), v_(v), loss_hist_(loss_hist) {
    TORCH_CHECK(world >= 1 && world <= 8, "world size 1..8 (one NVSwitch domain)");
    a_ = st::UpdateArgs{};
    a_.world = int(world); a_.rank = int(rank);
    for (int r = 0; r < world; ++r) {
      a_.grad[r] = reinterpret_cast<float*>(grad_ptrs[r]);
      a_.param[r] = reinterpret_cast<float*>(param_ptrs[r]);
      a_.wbf16[r] = wbf16_ptrs.empty() ? nullptr : reinterpret_cast<__nv_bfloat16*>(wbf16_ptrs[r]);
      a_.flags[r] = flag_ptrs.empty() ? nullptr : reinterpret_cast<uint32_t*>(flag_ptrs[r]);
    }
    a_.mc_grad = reinterpret_cast<float*>(mc_grad_ptr);
    a_.mc_param = reinterpret_cast<float*>(mc_param_ptr);
    a_.mc_wbf16 = reinterpret_cast<__nv_bfloat16*>(mc_wbf16_ptr);
    a_.m = m ? m->data_ptr<float>() : nullptr;
    a_.v = v ? v->data_ptr<float>() : nullptr;
    TORCH_CHECK(state.numel() * state.element_size() >= int64_t(sizeof(st::StepState)), "state tensor too small");
    a_.state = reinterpret_cast<st::StepState*>(state.data_ptr());
    a_.loss_hist = loss_hist ? loss_hist->data_ptr<float>() : nullptr;
    a_.max_iters = loss_hist ? loss_hist->numel() : 0;
    a_.host_mailbox = reinterpret_cast<volatile uint32_t*>(host_mailbox_ptr);
    TORCH_CHECK(segs.is_cuda() && segs.numel() * segs.element_size() % sizeof(st::ParamSegment) == 0);
    a_.segs = reinterpret_cast<const st::ParamSegment*>(segs.data_ptr());
    a_.n_segs = int(segs.numel() * segs.element_size() / sizeof(st::ParamSegment));
    a_.numel = numel; a_.tail = int(tail);
    a_.kind = int(kind); a_.lr = float(lr); a_.beta1 = float(beta1); a_.beta2 = float(beta2); a_.eps = float(eps);
    a_.weight_decay = float(weight_decay); a_.momentum = float(momentum); a_.dampening = float(dampening);
    a_.nesterov = nesterov; a_.maximize = maximize; a_.patience = int(patience); a_.zero_grads = zero_grads;
    a_.recv_local = reinterpret_cast<float*>(recv_local_ptr); a_.recv_stride = recv_stride;
  }
  void step() {
    if (a_.recv_local != nullptr) check(st::allreduce_push_update(a_, cur_stream()), "allreduce_push_update");
    else check(st::allreduce_update(a_, cur_stream()), "allreduce_update");
  }
  void publish() { check(st::publish_params(a_, cur_stream()), "publish_params"); }
  void set_lr(double lr) { a_.lr = float(lr); }

 private:
  st::UpdateArgs a_;
  Tensor state_, segs_;
  OptT m_, v_, loss_hist_;
};

void allreduce_sum(std::vector<int64_t> buf_ptrs, std::vector<int64_t> flag_ptrs, const Tensor& state, int64_t world,
                   int64_t rank, int64_t n, int64_t multicast_ptr) {
  float* bufs[8];
  uint32_t* flags[8];
  for (int r = 0; r < world; ++r) { bufs[r] = reinterpret_cast<float*>(buf_ptrs[r]); flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]); }
  check(st::allreduce_sum_f32(bufs, flags, reinterpret_cast<st::StepState*>(state.data_ptr()), int(world), int(rank), n,
                              reinterpret_cast<float*>(multicast_ptr), cur_stream()),
        "allreduce_sum");
}

// ------------------------------------------------------------------ hogwild over peer memory
class HogwildWorker {
 public:
  HogwildWorker(int64_t world, int64_t rank, int64_t stripes, std::vector<int64_t> param_ptrs, std::vector<int64_t> m_ptrs,
                std::vector<int64_t> v_ptrs, std::vector<int64_t> ctrl_ptrs, const Tensor& local_param, const Tensor& local_grad,
                const Tensor& loss_in, bool use_val_loss, int64_t numel, bool acquire_lock, int64_t kind, double lr, double beta1,
                double beta2, double eps, double weight_decay, double momentum, double dampening, bool nesterov, bool maximize,
                int64_t patience, int64_t window_len, int64_t host_mailbox_ptr, const OptT& local_state, const OptT& loss_hist,
                bool clear_after_push)
      : lp_(local_param), lg_(local_grad), li_(loss_in), ls_(local_state), lh_(loss_hist) {
    a_ = st::HogwildArgs{};
    a_.world = int(world); a_.rank = int(rank); a_.stripes = int(stripes);
    for (int r = 0; r < world; ++r) {
      a_.param[r] = reinterpret_cast<float*>(param_ptrs[r]);
      a_.m[r] = m_ptrs.empty() ? nullptr : reinterpret_cast<float*>(m_ptrs[r]);
      a_.v[r] = v_ptrs.empty() ? nullptr : reinterpret_cast<float*>(v_ptrs[r]);
      a_.ctrl[r] = reinterpret_cast<uint32_t*>(ctrl_ptrs[r]);
    }
    a_.local_param = local_param.data_ptr<float>(); a_.local_grad = local_grad.data_ptr<float>();
    a_.loss_in = loss_in.data_ptr<float>(); a_.use_val_loss = use_val_loss;
    a_.numel = numel; a_.acquire_lock = acquire_lock; a_.kind = int(kind);
    a_.lr = float(lr); a_.beta1 = float(beta1); a_.beta2 = float(beta2); a_.eps = float(eps);
    a_.weight_decay = float(weight_decay); a_.momentum = float(momentum); a_.dampening = float(dampening);
    a_.nesterov = nesterov; a_.maximize = maximize; a_.patience = int(patience); a_.window_len = int(window_len);
    a_.host_mailbox = reinterpret_cast<volatile uint32_t*>(host_mailbox_ptr);
    a_.local_state = local_state ? reinterpret_cast<st::StepState*>(local_state->data_ptr()) : nullptr;
    a_.loss_hist = loss_hist ? loss_hist->data_ptr<float>() : nullptr;
    a_.max_iters = loss_hist ? loss_hist->numel() : 0;
    a_.loss_slot_rw = clear_after_push ? loss_in.data_ptr<float>() : nullptr;
    a_.local_grad_rw = clear_after_push ? local_grad.data_ptr<float>() : nullptr;
    a_.dry_run = 0;
  }
  void pull() { check(st::hogwild_pull(a_, cur_stream()), "hogwild_pull"); }
  void push() { check(st::hogwild_push(a_, cur_stream()), "hogwild_push"); }
  void step() { push(); }
  void warm() { a_.dry_run = 1; push(); a_.dry_run = 0; }

 private:
  st::HogwildArgs a_;
  Tensor lp_, lg_, li_;
  OptT ls_, lh_;
};
int64_t hw_ctrl_words() { return st::HW_CTRL_WORDS; }
int64_t hw_max_stripes() { return st::HW_MAX_STRIPES; }


// ------------------------------------------------------------------ K-STEP: whole training steps in one persistent kernel
// (csrc/mlp_step.cu).  The task graph, tensor-map table and peer pointer table are built by runtime/mega_mlp.py.
py::bytes make_tmap_bytes(int64_t ptr, int64_t rows, int64_t cols, int64_t ld, int64_t box_rows) {
  CUtensorMap map;
  check(st::make_tmap_bf16_2d(&map, reinterpret_cast<const void*>(ptr), uint64_t(rows), uint64_t(cols), uint64_t(ld), uint32_t(box_rows)),
        "cuTensorMapEncodeTiled");
  return py::bytes(reinterpret_cast<const char*>(&map), sizeof(map));
}

class MlpStep {
 public:
  MlpStep(const Tensor& tasks, const Tensor& n_tasks, const Tensor& tmaps, const Tensor& peers, const Tensor& counters,
          int64_t flags_local_ptr, int64_t world, int64_t rank, int64_t grad_ptr, int64_t param_ptr, int64_t wbf16_ptr,
          int64_t mc_grad_ptr, int64_t mc_param_ptr, int64_t mc_wbf16_ptr, const OptT& m, const OptT& v, int64_t shard_lo,
          const Tensor& segs, const Tensor& state, const OptT& loss_hist, int64_t mailbox_ptr, int64_t kind, double lr, double beta1,
          double beta2, double eps, double weight_decay, double momentum, double dampening, bool nesterov, bool maximize,
          int64_t patience, int64_t decide_cnt, int64_t n_mt, int64_t loss_off, int64_t window_rows, int64_t n_windows,
          const Tensor& error)
      : tasks_(tasks), n_tasks_(n_tasks), tmaps_(tmaps), peers_(peers), counters_(counters), segs_(segs), state_(state),
        error_(error), m_(m), v_(v), loss_hist_(loss_hist) {
    TORCH_CHECK(tasks.is_cuda() && n_tasks.is_cuda() && tmaps.is_cuda() && peers.is_cuda() && counters.is_cuda() && segs.is_cuda() && error.is_cuda());
    TORCH_CHECK(n_tasks.scalar_type() == at::kInt && counters.scalar_type() == at::kInt && error.scalar_type() == at::kInt && error.numel() >= 4);
    grid_ = int(n_tasks.numel());
    TORCH_CHECK(int64_t(tasks.numel() * tasks.element_size()) == int64_t(grid_) * st::MS_MAX_TASKS * int64_t(sizeof(st::MsTask)), "task table size");
    TORCH_CHECK(int64_t(peers.numel() * peers.element_size()) == int64_t(sizeof(st::MsPeerTable)), "peer table size");
    TORCH_CHECK((reinterpret_cast<uintptr_t>(tmaps.data_ptr()) & 63) == 0 && (reinterpret_cast<uintptr_t>(tasks.data_ptr()) & 15) == 0);
    p_ = st::MsParams{};
    p_.tasks = reinterpret_cast<const st::MsTask*>(tasks.data_ptr());
    p_.n_tasks = n_tasks.data_ptr<int32_t>();
    p_.tmaps = reinterpret_cast<const CUtensorMap*>(tmaps.data_ptr());
    p_.peers = reinterpret_cast<const st::MsPeerTable*>(peers.data_ptr());
    p_.counters = reinterpret_cast<uint32_t*>(counters.data_ptr<int32_t>());
    p_.flags_local = reinterpret_cast<uint32_t*>(flags_local_ptr);
    p_.world = int(world); p_.rank = int(rank);
    p_.grad_local = reinterpret_cast<float*>(grad_ptr);
    p_.param_local = reinterpret_cast<float*>(param_ptr);
    p_.wbf16_local = reinterpret_cast<__nv_bfloat16*>(wbf16_ptr);
    p_.mc_grad = reinterpret_cast<float*>(mc_grad_ptr);
    p_.mc_param = reinterpret_cast<float*>(mc_param_ptr);
    p_.mc_wbf16 = reinterpret_cast<__nv_bfloat16*>(mc_wbf16_ptr);
    p_.m = m ? m->data_ptr<float>() : nullptr;
    p_.v = v ? v->data_ptr<float>() : nullptr;
    p_.shard_lo = shard_lo;
    p_.segs = reinterpret_cast<const st::ParamSegment*>(segs.data_ptr());
    TORCH_CHECK(state.numel() * state.element_size() >= int64_t(sizeof(st::StepState)));
    p_.state = reinterpret_cast<st::StepState*>(state.data_ptr());
    p_.loss_hist = loss_hist ? loss_hist->data_ptr<float>() : nullptr;
    p_.max_iters = loss_hist ? loss_hist->numel() : 0;
    p_.mailbox = reinterpret_cast<volatile uint32_t*>(mailbox_ptr);
    p_.kind = int(kind); p_.lr = float(lr); p_.beta1 = float(beta1); p_.beta2 = float(beta2); p_.eps = float(eps);
    p_.weight_decay = float(weight_decay); p_.momentum = float(momentum); p_.dampening = float(dampening);
    p_.nesterov = nesterov; p_.maximize = maximize; p_.patience = int(patience); p_.decide_cnt = int(decide_cnt);
    p_.n_mt = int(n_mt); p_.loss_off = loss_off; p_.window_rows = window_rows; p_.n_windows = n_windows > 0 ? n_windows : 1;
    p_.error = reinterpret_cast<uint32_t*>(error.data_ptr<int32_t>());
  }
  // cross-GPU dependencies as counters in the symmetric flag block (wait kind 3); mc_flags_ptr: its multicast alias or 0
  void set_xcnt(int64_t mc_flags_ptr) {
    p_.xcnt = 1;
    p_.mc_flags = reinterpret_cast<uint32_t*>(mc_flags_ptr);
  }
  void set_sys_fence_all(bool on) { p_.sys_fence_all = on ? 1 : 0; }
  void set_prof_kb(const OptT& prof) {
    prof_kb_ = prof;
    if (prof) {
      TORCH_CHECK(prof->is_cuda() && prof->scalar_type() == at::kLong && prof->numel() >= int64_t(grid_) * st::MS_MAX_TASKS * 64);
      p_.prof_kb = reinterpret_cast<long long*>(prof->data_ptr());
    } else {
      p_.prof_kb = nullptr;
    }
  }
  void set_prof(const OptT& prof) {
    prof_ = prof;
    if (prof) {
      TORCH_CHECK(prof->is_cuda() && prof->scalar_type() == at::kLong && prof->numel() >= int64_t(grid_) * st::MS_MAX_TASKS * 8);
      p_.prof = reinterpret_cast<long long*>(prof->data_ptr());
    } else {
      p_.prof = nullptr;
    }
  }
  // n_steps whole training steps in one launch on the current stream
  void run(int64_t iter0, int64_t step0, int64_t epoch0, int64_t n_steps) {
    p_.iter0 = iter0; p_.step0 = step0; p_.epoch0 = uint32_t(epoch0); p_.n_steps = int(n_steps);
    check(st::mlp_step_launch(p_, grid_, cur_stream()), "mlp_step_launch");
  }
  // `iters` steps, `per_launch` steps per kernel; returns the steps launched (stops early once the device latched a stop)
  int64_t run_many(int64_t iter0, int64_t step0, int64_t epoch0, int64_t iters, int64_t per_launch, int64_t mailbox_ptr) {
    volatile uint32_t* mb = reinterpret_cast<volatile uint32_t*>(mailbox_ptr);
    int64_t done = 0;
    if (per_launch < 1) per_launch = 1;
    while (done < iters) {
      const int64_t n = std::min(per_launch, iters - done);
      run(iter0 + done, step0 + done, epoch0 + done, n);
      done += n;
      if (mb != nullptr && mb[0] != 0) break;
    }
    return done;
  }
  int64_t grid() const { return grid_; }

 private:
  st::MsParams p_;
  int grid_ = 0;
  Tensor tasks_, n_tasks_, tmaps_, peers_, counters_, segs_, state_, error_;
  OptT m_, v_, loss_hist_, prof_, prof_kb_;
};
int64_t ms_task_bytes() { return int64_t(sizeof(st::MsTask)); }
int64_t ms_max_tasks() { return st::MS_MAX_TASKS; }
int64_t ms_max_flags() { return st::MS_MAX_FLAGS; }
int64_t ms_smem_bytes() { return st::ms_smem_bytes(); }

py::bytes pack_segments(const std::vector<std::vector<int64_t>>& rows) {
  std::vector<st::ParamSegment> segs(rows.size());
  for (size_t i = 0; i < rows.size(); ++i) {
    const auto& r = rows[i];
    TORCH_CHECK(r.size() == 8, "segment = [offset, rows, cols, w_off, w_ld, wt_off, wt_ld, publish_f32]");
    segs[i].offset = r[0]; segs[i].rows = int32_t(r[1]); segs[i].cols = int32_t(r[2]); segs[i].w_off = r[3];
    segs[i].w_ld = int32_t(r[4]); segs[i].wt_off = r[5]; segs[i].wt_ld = int32_t(r[6]); segs[i].publish_f32 = int32_t(r[7]);
  }
  return py::bytes(reinterpret_cast<const char*>(segs.data()), segs.size() * sizeof(st::ParamSegment));
}

// ------------------------------------------------------------------ native step runner
// Captures whatever the Python side launches on the current stream between
// begin_capture()/end_capture() into a CUDA graph, then replays it `iters` times
// from C++ – no Python in the training loop.  Optional input streaming: before
// iteration i the batch [i*B, (i+1)*B) (mod n) of a pinned host matrix is copied
// into staging slot i % ring on a copy stream; the graph of iteration i waits for
// that copy, the copy of iteration i+ring waits for graph i (slot reuse).
class GraphRunner {
 public:
  GraphRunner() = default;
  ~GraphRunner() {
    if (exec_) cudaGraphExecDestroy(exec_);
    if (graph_) cudaGraphDestroy(graph_);
    if (tail_exec_) cudaGraphExecDestroy(tail_exec_);
    if (tail_graph_) cudaGraphDestroy(tail_graph_);
    for (auto e : copied_) cudaEventDestroy(e);
    for (auto e : consumed_) cudaEventDestroy(e);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
  }
  void begin_capture() {
    stream_ = cur_stream();
    TORCH_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal) == cudaSuccess, "begin capture failed");
  }
  void end_capture() {
    TORCH_CHECK(cudaStreamEndCapture(stream_, &graph_) == cudaSuccess && graph_ != nullptr, "end capture failed");
    TORCH_CHECK(cudaGraphInstantiate(&exec_, graph_, 0) == cudaSuccess, "graph instantiate failed");
    size_t n = 0;
    cudaGraphGetNodes(graph_, nullptr, &n);
    n_nodes_ = int64_t(n);
  }
  int64_t num_nodes() const { return n_nodes_; }

  void setup_streaming(int64_t host_x_ptr, int64_t host_y_ptr, int64_t n_rows, int64_t x_row_bytes, int64_t y_row_bytes,
                       int64_t dev_x_ptr, int64_t dev_y_ptr, int64_t batch, int64_t ring) {
    hx_ = reinterpret_cast<const char*>(host_x_ptr); hy_ = reinterpret_cast<const char*>(host_y_ptr);
    dx_ = reinterpret_cast<char*>(dev_x_ptr); dy_ = reinterpret_cast<char*>(dev_y_ptr);
    n_rows_ = n_rows; xrb_ = x_row_bytes; yrb_ = y_row_bytes; batch_ = batch; ring_ = ring;
    TORCH_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking) == cudaSuccess);
    copied_.resize(ring); consumed_.resize(ring);
    for (int i = 0; i < ring; ++i) {
      cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&consumed_[i], cudaEventDisableTiming);
    }
    streaming_ = true;
  }

  // The graph captured next (begin_capture/end_capture) runs `spl` training steps per launch (a K-STEP kernel with n_steps =
  // spl); the previously captured one-step graph stays as the tail graph for iteration counts that are no multiple of spl.
  void keep_as_tail_and_expect(int64_t spl) {
    TORCH_CHECK(exec_ != nullptr && spl >= 1, "capture the one-step graph first");
    tail_exec_ = exec_; tail_graph_ = graph_;
    exec_ = nullptr; graph_ = nullptr;
    spl_ = spl;
  }

  // returns number of iterations launched; stops early when the device latched a stop in the mailbox
  int64_t run(int64_t first_iter, int64_t iters, int64_t mailbox_ptr, int64_t check_every) {
    TORCH_CHECK(exec_ != nullptr, "no captured graph");
    volatile uint32_t* mb = reinterpret_cast<volatile uint32_t*>(mailbox_ptr);
    cudaStream_t s = cur_stream();
    int64_t done = 0;
    const int64_t windows = streaming_ ? (n_rows_ / batch_) : 0;
    int64_t k = 0;
    while (k < iters) {
      // steps of this launch: spl_ with the main graph, 1 with the tail graph
      const bool big = spl_ > 1 && iters - k >= spl_;
      const int64_t n = big ? spl_ : 1;
      cudaGraphExec_t ex = (spl_ > 1 && !big) ? tail_exec_ : exec_;
      if (streaming_) {
        int last_slot = 0;
        for (int64_t j = 0; j < n; ++j) {
          const int64_t i = first_iter + k + j;
          const int slot = int(i % ring_);
          if (i >= ring_) cudaStreamWaitEvent(copy_stream_, consumed_[slot], 0);
          const int64_t row0 = (i % windows) * batch_;
          cudaMemcpyAsync(dx_ + size_t(slot) * batch_ * xrb_, hx_ + size_t(row0) * xrb_, size_t(batch_) * xrb_, cudaMemcpyHostToDevice, copy_stream_);
          if (hy_ != nullptr)
            cudaMemcpyAsync(dy_ + size_t(slot) * batch_ * yrb_, hy_ + size_t(row0) * yrb_, size_t(batch_) * yrb_, cudaMemcpyHostToDevice, copy_stream_);
          h2d_bytes_ += uint64_t(batch_) * uint64_t(xrb_ + (hy_ ? yrb_ : 0));
          last_slot = slot;
        }
        cudaEventRecord(copied_[last_slot], copy_stream_);       // the copy stream is in order: the last batch implies the others
        cudaStreamWaitEvent(s, copied_[last_slot], 0);
      }
      TORCH_CHECK(cudaGraphLaunch(ex, s) == cudaSuccess, "graph launch failed");
      if (streaming_)
        for (int64_t j = 0; j < n; ++j) cudaEventRecord(consumed_[int((first_iter + k + j) % ring_)], s);
      done += n;
      k += n;
      if (mb != nullptr && check_every > 0 && mb[0] != 0) break;
    }
    return done;
  }
  int64_t h2d_bytes() const { return int64_t(h2d_bytes_); }

 private:
  cudaStream_t stream_ = nullptr, copy_stream_ = nullptr;
  cudaGraph_t graph_ = nullptr, tail_graph_ = nullptr;
  cudaGraphExec_t exec_ = nullptr, tail_exec_ = nullptr;
  int64_t spl_ = 1;
  int64_t n_nodes_ = 0;
  bool streaming_ = false;
  const char *hx_ = nullptr, *hy_ = nullptr;
  char *dx_ = nullptr, *dy_ = nullptr;
  int64_t n_rows_ = 0, xrb_ = 0, yrb_ = 0, batch_ = 0, ring_ = 1;
  std::vector<cudaEvent_t> copied_, consumed_;
  uint64_t h2d_bytes_ = 0;
};

void set_persistent_mode(int64_t m) { st::g_persistent_mode = int(m); }
void set_pair_mode(int64_t m) { st::g_pair_mode = int(m); }
void set_pdl(bool on) { st::g_use_pdl = on; }
bool get_pdl() { return st::g_use_pdl; }
int64_t step_state_bytes() { return int64_t(sizeof(st::StepState)); }
int64_t flag_words() { return st::AR_FLAG_WORDS; }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "sparktorch_b200 native kernels (sm_100a)";
  m.def("conv2d_fwd_implicit", [](const Tensor& x, const Tensor& w, Tensor& out, const c10::optional<Tensor>& bias, bool relu, int64_t R,
                                  int64_t S, int64_t stride, int64_t pad, int64_t dil) {
    TORCH_CHECK(x.is_cuda() && w.is_cuda() && out.is_cuda() && x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 &&
                out.scalar_type() == at::kBFloat16 && x.dim() == 4 && x.is_contiguous() && w.dim() == 2 && w.stride(1) == 1 && out.dim() == 2 &&
                out.stride(1) == 1, "x: [N, H, W, C] bf16 contiguous; w: [Cout, R*S*C] bf16; out: [N*P*Q, Cout] bf16");
    TORCH_CHECK(w.size(1) == R * S * x.size(3) && out.size(1) == w.size(0));
    st::GemmEpilogue ep;
    ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out.data_ptr());
    ep.ld_out = out.stride(0);
    ep.relu = relu ? 1 : 0;
    if (bias.has_value()) { TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == at::kFloat && bias->numel() >= w.size(0)); ep.bias = bias->data_ptr<float>(); }
    return int64_t(st::conv2d_fwd_implicit(x.data_ptr(), int(x.size(0)), int(x.size(1)), int(x.size(2)), int(x.size(3)), w.data_ptr(), w.stride(0),
                                           int(w.size(0)), int(R), int(S), int(stride), int(pad), int(dil), &ep, 0, cur_stream()));
  }, "forward convolution as an implicit GEMM (TMA im2col loads, tcgen05 MMA): returns 0, or -2 when the shape is not handled");
  m.def("gemm_bf16_2cta", [](const Tensor& A, const Tensor& B, Tensor& Cout, const c10::optional<Tensor>& bias, double alpha, bool relu) {
    TORCH_CHECK(A.is_cuda() && B.is_cuda() && Cout.is_cuda() && A.scalar_type() == at::kBFloat16 && B.scalar_type() == at::kBFloat16 &&
                Cout.scalar_type() == at::kBFloat16 && A.dim() == 2 && B.dim() == 2 && Cout.dim() == 2);
    TORCH_CHECK(A.stride(1) == 1 && B.stride(1) == 1 && Cout.stride(1) == 1 && A.size(1) == B.size(1) && Cout.size(0) == A.size(0) && Cout.size(1) == B.size(0));
    const float* bp = nullptr;
    if (bias.has_value()) { TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == at::kFloat && bias->numel() >= B.size(0)); bp = bias->data_ptr<float>(); }
    return st::gemm_bf16_2cta(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cout.data_ptr(), Cout.stride(0), int(A.size(0)), int(B.size(0)),
                              int(A.size(1)), bp, float(alpha), relu ? 1 : 0, cur_stream());
  }, "bf16 GEMM on CTA pairs (cta_group::2): C = act(alpha A B^T + bias); returns 0, or -2 when the shape is left to gemm_bf16");
  m.def("gemm_bf16", &gemm_bf16, "tcgen05 bf16 GEMM with fused epilogue", py::arg("A"), py::arg("a_mn"), py::arg("B"),
        py::arg("b_mn"), py::arg("M"), py::arg("N"), py::arg("K"), py::arg("alpha") = 1.0, py::arg("bias") = py::none(),
        py::arg("relu") = false, py::arg("mask_src") = py::none(), py::arg("out_bf16") = py::none(),
        py::arg("out_t_bf16") = py::none(), py::arg("out_f32") = py::none(), py::arg("out_t_f32") = py::none(),
        py::arg("atomic_f32") = false, py::arg("colsum") = py::none(), py::arg("block_n") = 0, py::arg("split_k") = 1,
        py::arg("loss_kind") = 0, py::arg("labels") = py::none(), py::arg("target") = py::none(),
        py::arg("loss_slot") = py::none(), py::arg("loss_scale") = 1.0, py::arg("win_state") = py::none(),
        py::arg("window_rows") = 0, py::arg("n_windows") = 1, py::arg("win_flags") = 0,
        py::arg("prof") = py::none(), py::arg("peer_table") = py::none(), py::arg("peer_shard_elems") = 0,
        py::arg("peer_w_off") = 0, py::arg("peer_bias_off") = -1, py::arg("peer_fill_local") = false,
        py::arg("push_table") = py::none(), py::arg("push_me") = 0, py::arg("push_mask") = 0,
        py::arg("push_state") = py::none(), py::arg("colsumsq") = py::none());
  m.def("softmax_ce", &softmax_ce, py::arg("logits"), py::arg("labels"), py::arg("C"), py::arg("loss_slot"), py::arg("dz"), py::arg("dzt"),
        py::arg("db"), py::arg("pad_cols"), py::arg("win_state") = py::none(), py::arg("window_rows") = 0, py::arg("n_windows") = 1);
  m.def("mse_loss", &mse_loss, py::arg("pred"), py::arg("target"), py::arg("D"), py::arg("loss_slot"), py::arg("dz"), py::arg("dzt"),
        py::arg("db"), py::arg("pad_cols"), py::arg("win_state") = py::none(), py::arg("window_rows") = 0, py::arg("n_windows") = 1);
  m.def("gather_batch", &gather_batch);
  m.def("transpose_bf16", &transpose_bf16);
  m.def("im2col_nhwc", &im2col_nhwc, py::arg("x"), py::arg("B"), py::arg("H"), py::arg("W"), py::arg("C"), py::arg("sb"),
        py::arg("sh"), py::arg("sw"), py::arg("sc"), py::arg("kh"), py::arg("kw"), py::arg("stride"), py::arg("pad"),
        py::arg("col"), py::arg("state") = py::none(), py::arg("window_rows") = 0, py::arg("n_windows") = 1);
  m.def("flip_conv_weight", &flip_conv_weight);
  m.def("maxpool_dropout", &maxpool_dropout, py::arg("backward"), py::arg("x_or_dx"), py::arg("ld_x"), py::arg("y_or_dy"),
        py::arg("idx"), py::arg("B"), py::arg("H"), py::arg("W"), py::arg("C"), py::arg("k"), py::arg("p_drop") = 0.0,
        py::arg("train") = false, py::arg("seed") = 0, py::arg("state") = py::none(), py::arg("y_nhwc") = false,
        py::arg("relu_x") = py::none(), py::arg("db") = py::none());
  m.def("im2col_cl", &im2col_cl);
  m.def("flip_conv_weight_cl", &flip_conv_weight_cl);
  m.def("bn_stats", &bn_stats);
  m.def("bn_finalize", &bn_finalize, py::arg("sums"), py::arg("M"), py::arg("C"), py::arg("eps"), py::arg("momentum"),
        py::arg("running_mean") = py::none(), py::arg("running_var") = py::none(), py::arg("stats"),
        py::arg("from_running") = false, py::arg("update_running") = true);
  m.def("bn_apply", &bn_apply, py::arg("y"), py::arg("M"), py::arg("C"), py::arg("stats"), py::arg("gamma"), py::arg("beta"),
        py::arg("z"), py::arg("relu") = false, py::arg("residual") = py::none(), py::arg("sums") = py::none(), py::arg("eps") = 1e-5,
        py::arg("momentum") = 0.1, py::arg("running_mean") = py::none(), py::arg("running_var") = py::none());
  m.def("bn_backward", &bn_backward, py::arg("reduce"), py::arg("y"), py::arg("M"), py::arg("C"), py::arg("stats"), py::arg("gamma"),
        py::arg("z"), py::arg("relu"), py::arg("dz_a"), py::arg("dz_b") = py::none(), py::arg("dgamma"), py::arg("dbeta"),
        py::arg("dy") = py::none(), py::arg("g_out") = py::none());
  m.def("maxpool_cl", &maxpool_cl, py::arg("backward"), py::arg("x_or_dx"), py::arg("y_or_dy"), py::arg("idx"), py::arg("B"),
        py::arg("H"), py::arg("W"), py::arg("C"), py::arg("k"), py::arg("s"), py::arg("p"), py::arg("dy_b") = py::none());
  m.def("avgpool_cl", &avgpool_cl);
  m.def("allreduce_sum", &allreduce_sum, py::arg("buf_ptrs"), py::arg("flag_ptrs"), py::arg("state"), py::arg("world"),
        py::arg("rank"), py::arg("n"), py::arg("multicast_ptr") = 0);
  m.def("pack_segments", &pack_segments);
  m.def("step_state_bytes", &step_state_bytes);
  m.def("set_pdl", &set_pdl);
  m.def("set_persistent_mode", &set_persistent_mode);
  m.def("set_pair_mode", &set_pair_mode, "0: gemm_bf16 never hands big plain products to the CTA-pair kernel");
  m.def("get_pdl", &get_pdl);
  m.def("flag_words", &flag_words);
  py::class_<Updater>(m, "Updater")
      .def(py::init<int64_t, int64_t, std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>,
                    int64_t, const OptT&, const OptT&, const Tensor&, const OptT&, int64_t, const Tensor&, int64_t, int64_t,
                    int64_t, double, double, double, double, double, double, double, bool, bool, int64_t, bool, int64_t, int64_t,
                    int64_t, int64_t>())
      .def("step", &Updater::step)
      .def("publish", &Updater::publish)
      .def("set_lr", &Updater::set_lr);
  m.def("hw_ctrl_words", &hw_ctrl_words);
  m.def("hw_max_stripes", &hw_max_stripes);
  py::class_<HogwildWorker>(m, "HogwildWorker")
      .def(py::init<int64_t, int64_t, int64_t, std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>,
                    std::vector<int64_t>, const Tensor&, const Tensor&, const Tensor&, bool, int64_t, bool, int64_t, double,
                    double, double, double, double, double, double, bool, bool, int64_t, int64_t, int64_t, const OptT&,
                    const OptT&, bool>())
      .def("pull", &HogwildWorker::pull)
      .def("push", &HogwildWorker::push)
      .def("step", &HogwildWorker::step)
      .def("warm", &HogwildWorker::warm);
  m.def("make_tmap_bytes", &make_tmap_bytes);
  m.def("ms_task_bytes", &ms_task_bytes);
  m.def("ms_max_tasks", &ms_max_tasks);
  m.def("ms_max_flags", &ms_max_flags);
  m.def("ms_smem_bytes", &ms_smem_bytes);
  py::class_<MlpStep>(m, "MlpStep")
      .def(py::init<const Tensor&, const Tensor&, const Tensor&, const Tensor&, const Tensor&, int64_t, int64_t, int64_t, int64_t,
                    int64_t, int64_t, int64_t, int64_t, int64_t, const OptT&, const OptT&, int64_t, const Tensor&, const Tensor&,
                    const OptT&, int64_t, int64_t, double, double, double, double, double, double, double, bool, bool, int64_t,
                    int64_t, int64_t, int64_t, int64_t, int64_t, const Tensor&>())
      .def("set_prof", &MlpStep::set_prof)
      .def("set_xcnt", &MlpStep::set_xcnt)
      .def("set_sys_fence_all", &MlpStep::set_sys_fence_all)
      .def("set_prof_kb", &MlpStep::set_prof_kb)
      .def("run", &MlpStep::run)
      .def("run_many", &MlpStep::run_many, py::call_guard<py::gil_scoped_release>())
      .def("grid", &MlpStep::grid);
  py::class_<GraphRunner>(m, "GraphRunner")
      .def(py::init<>())
      .def("begin_capture", &GraphRunner::begin_capture)
      .def("end_capture", &GraphRunner::end_capture)
      .def("num_nodes", &GraphRunner::num_nodes)
      .def("setup_streaming", &GraphRunner::setup_streaming)
      .def("keep_as_tail_and_expect", &GraphRunner::keep_as_tail_and_expect)
      .def("run", &GraphRunner::run, py::call_guard<py::gil_scoped_release>())
      .def("h2d_bytes", &GraphRunner::h2d_bytes);
}
