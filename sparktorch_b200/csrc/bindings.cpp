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
               int64_t push_mask, const OptT& push_state) {
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

void softmax_ce(const Tensor& logits, const Tensor& labels, int64_t C, const Tensor& loss_slot, const OptT& dz,
                const OptT& dzt, const OptT& db, int64_t pad_cols) {
  req_2d(logits, at::kFloat, "logits");
  TORCH_CHECK(labels.scalar_type() == at::kInt && loss_slot.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(logits.device());
  check(st::softmax_ce(logits.data_ptr<float>(), logits.stride(0), labels.data_ptr<int32_t>(), int(logits.size(0)), int(C),
                       loss_slot.data_ptr<float>(), dz ? bf16pm(*dz) : nullptr, dz ? dz->stride(0) : 0,
                       dzt ? bf16pm(*dzt) : nullptr, dzt ? dzt->stride(0) : 0, db ? db->data_ptr<float>() : nullptr,
                       int(pad_cols), cur_stream()),
        "softmax_ce");
}

void mse_loss(const Tensor& pred, const Tensor& target, int64_t D, const Tensor& loss_slot, const OptT& dz, const OptT& dzt,
              const OptT& db, int64_t pad_cols) {
  req_2d(pred, at::kFloat, "pred");
  req_2d(target, at::kFloat, "target");
  c10::cuda::CUDAGuard guard(pred.device());
  check(st::mse_loss(pred.data_ptr<float>(), pred.stride(0), target.data_ptr<float>(), target.stride(0), int(pred.size(0)),
                     int(D), loss_slot.data_ptr<float>(), dz ? bf16pm(*dz) : nullptr, dz ? dz->stride(0) : 0,
                     dzt ? bf16pm(*dzt) : nullptr, dzt ? dzt->stride(0) : 0, db ? db->data_ptr<float>() : nullptr, int(pad_cols),
                     cur_stream()),
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

// ------------------------------------------------------------------ fused all-reduce + optimizer
class Updater {
 public:
  Updater(int64_t world, int64_t rank, std::vector<int64_t> grad_ptrs, std::vector<int64_t> param_ptrs,
          std::vector<int64_t> wbf16_ptrs, std::vector<int64_t> flag_ptrs, int64_t mc_grad_ptr, const OptT& m, const OptT& v,
          const Tensor& state, const OptT& loss_hist, int64_t host_mailbox_ptr, const Tensor& segs, int64_t numel,
          int64_t tail, int64_t kind, double lr, double beta1, double beta2, double eps, double weight_decay, double momentum,
          double dampening, bool nesterov, bool maximize, int64_t patience, bool zero_grads, int64_t recv_local_ptr,
          int64_t recv_stride)
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
                   int64_t rank, int64_t n) {
  float* bufs[8];
  uint32_t* flags[8];
  for (int r = 0; r < world; ++r) { bufs[r] = reinterpret_cast<float*>(buf_ptrs[r]); flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]); }
  check(st::allreduce_sum_f32(bufs, flags, reinterpret_cast<st::StepState*>(state.data_ptr()), int(world), int(rank), n, 1, cur_stream()),
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

  // returns number of iterations launched; stops early when the device latched a stop in the mailbox
  int64_t run(int64_t first_iter, int64_t iters, int64_t mailbox_ptr, int64_t check_every) {
    TORCH_CHECK(exec_ != nullptr, "no captured graph");
    volatile uint32_t* mb = reinterpret_cast<volatile uint32_t*>(mailbox_ptr);
    cudaStream_t s = cur_stream();
    int64_t done = 0;
    const int64_t windows = streaming_ ? (n_rows_ / batch_) : 0;
    for (int64_t k = 0; k < iters; ++k) {
      const int64_t i = first_iter + k;
      if (streaming_) {
        const int slot = int(i % ring_);
        if (i >= ring_) cudaStreamWaitEvent(copy_stream_, consumed_[slot], 0);
        const int64_t row0 = (i % windows) * batch_;
        cudaMemcpyAsync(dx_ + size_t(slot) * batch_ * xrb_, hx_ + size_t(row0) * xrb_, size_t(batch_) * xrb_, cudaMemcpyHostToDevice, copy_stream_);
        if (hy_ != nullptr)
          cudaMemcpyAsync(dy_ + size_t(slot) * batch_ * yrb_, hy_ + size_t(row0) * yrb_, size_t(batch_) * yrb_, cudaMemcpyHostToDevice, copy_stream_);
        cudaEventRecord(copied_[slot], copy_stream_);
        cudaStreamWaitEvent(s, copied_[slot], 0);
        h2d_bytes_ += uint64_t(batch_) * uint64_t(xrb_ + (hy_ ? yrb_ : 0));
      }
      TORCH_CHECK(cudaGraphLaunch(exec_, s) == cudaSuccess, "graph launch failed");
      if (streaming_) cudaEventRecord(consumed_[int(i % ring_)], s);
      ++done;
      if (mb != nullptr && check_every > 0 && (k % check_every) == check_every - 1 && mb[0] != 0) break;
    }
    return done;
  }
  int64_t h2d_bytes() const { return int64_t(h2d_bytes_); }

 private:
  cudaStream_t stream_ = nullptr, copy_stream_ = nullptr;
  cudaGraph_t graph_ = nullptr;
  cudaGraphExec_t exec_ = nullptr;
  int64_t n_nodes_ = 0;
  bool streaming_ = false;
  const char *hx_ = nullptr, *hy_ = nullptr;
  char *dx_ = nullptr, *dy_ = nullptr;
  int64_t n_rows_ = 0, xrb_ = 0, yrb_ = 0, batch_ = 0, ring_ = 1;
  std::vector<cudaEvent_t> copied_, consumed_;
  uint64_t h2d_bytes_ = 0;
};

void set_persistent_mode(int64_t m) { st::g_persistent_mode = int(m); }
void set_pdl(bool on) { st::g_use_pdl = on; }
bool get_pdl() { return st::g_use_pdl; }
int64_t step_state_bytes() { return int64_t(sizeof(st::StepState)); }
int64_t flag_words() { return st::AR_FLAG_WORDS; }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "sparktorch_b200 native kernels (sm_100a)";
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
        py::arg("push_state") = py::none());
  m.def("softmax_ce", &softmax_ce);
  m.def("mse_loss", &mse_loss);
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
  m.def("allreduce_sum", &allreduce_sum);
  m.def("pack_segments", &pack_segments);
  m.def("step_state_bytes", &step_state_bytes);
  m.def("set_pdl", &set_pdl);
  m.def("set_persistent_mode", &set_persistent_mode);
  m.def("get_pdl", &get_pdl);
  m.def("flag_words", &flag_words);
  py::class_<Updater>(m, "Updater")
      .def(py::init<int64_t, int64_t, std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>,
                    int64_t, const OptT&, const OptT&, const Tensor&, const OptT&, int64_t, const Tensor&, int64_t, int64_t,
                    int64_t, double, double, double, double, double, double, double, bool, bool, int64_t, bool, int64_t, int64_t>())
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
  py::class_<GraphRunner>(m, "GraphRunner")
      .def(py::init<>())
      .def("begin_capture", &GraphRunner::begin_capture)
      .def("end_capture", &GraphRunner::end_capture)
      .def("num_nodes", &GraphRunner::num_nodes)
      .def("setup_streaming", &GraphRunner::setup_streaming)
      .def("run", &GraphRunner::run, py::call_guard<py::gil_scoped_release>())
      .def("h2d_bytes", &GraphRunner::h2d_bytes);
}
