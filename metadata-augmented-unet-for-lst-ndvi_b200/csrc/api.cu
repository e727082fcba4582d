// api.cu -- the extern "C" surface declared in include/mau_b200.h
#include <algorithm>
#include <cstring>
#include <vector>
#include <new>
#include "plan.h"

using namespace mau;

struct mau_plan {
  Plan impl;
};

extern "C" {

const char* mau_last_error(void) { return last_error().c_str(); }
int mau_version(void) { return 100; }
int64_t mau_launch_count(void) { return (int64_t)g_launches.load(); }
int mau_set_sm_reserve(int n_sms) { set_sm_reserve(n_sms); return 0; }

int mau_plan_create(const mau_config* cfg, mau_plan** out) {
  if (!cfg || !out) return fail("mau_plan_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available: this library has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (%d devices)", cfg->device, ndev);
  MAU_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  MAU_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail("sm_100a kernels only: device %d is sm_%d%d", cfg->device, prop.major, prop.minor);
  mau_plan* p = new (std::nothrow) mau_plan();
  if (!p) return fail("out of host memory");
  p->impl.cfg = *cfg;
  int rc = p->impl.build();
  if (rc) { delete p; return rc; }
  *out = p;
  return 0;
}

int mau_plan_destroy(mau_plan* plan) {
  delete plan;
  return 0;
}

size_t mau_plan_workspace_bytes(const mau_plan* plan) { return plan ? plan->impl.ws_bytes : 0; }
int mau_plan_num_state(const mau_plan* plan) { return plan ? (int)plan->impl.state.size() : 0; }

int mau_plan_state_info(const mau_plan* plan, int i, int64_t* numel, int* role) {
  if (!plan || i < 0 || i >= (int)plan->impl.state.size()) return fail("state index out of range");
  if (numel) *numel = plan->impl.state[i].numel;
  if (role) *role = plan->impl.state[i].role;
  return 0;
}

int mau_plan_describe_config(const mau_config* cfg, char* buf, size_t buflen) {
  if (!cfg || !buf) return fail("null argument");
  Plan p;
  p.cfg = *cfg;
  p.dry = true;
  int rc = p.build();
  if (rc) return rc;
  if (p.describe_json.size() + 1 > buflen) return fail("describe buffer too small (%zu needed)", p.describe_json.size() + 1);
  memcpy(buf, p.describe_json.c_str(), p.describe_json.size() + 1);
  return 0;
}

int mau_plan_flops(const mau_plan* plan, double* fwd_flops, double* bwd_flops) {
  if (!plan) return fail("null plan");
  if (fwd_flops) *fwd_flops = plan->impl.fwd_flops;
  if (bwd_flops) *bwd_flops = plan->impl.bwd_flops;
  return 0;
}

int mau_plan_exec_flops(const mau_plan* plan, double* conv_flops_per_forward, double* conv_flops_per_backward) {
  if (!plan || !conv_flops_per_forward) return fail("null argument");
  *conv_flops_per_forward = plan->impl.exec_flops;
  if (conv_flops_per_backward) *conv_flops_per_backward = plan->impl.exec_bwd_flops;
  return 0;
}

int mau_plan_forward(mau_plan* plan, void* const* state_dev, const float* maps_dev, const float* temp_series_dev,
                     const float* metadata_dev, float* out_dev, void* stream) {
  if (!plan || !state_dev || !maps_dev || !out_dev) return fail("mau_plan_forward: null argument");
  Ctx c;
  c.state = state_dev; c.maps = maps_dev; c.series = temp_series_dev; c.md = metadata_dev; c.out = out_dev;
  c.st = static_cast<cudaStream_t>(stream);
  plan->impl.last_state.assign(state_dev, state_dev + plan->impl.state.size());
  plan->impl.last_series = temp_series_dev; plan->impl.last_md = metadata_dev;
  return plan->impl.run_forward(c);
}

int mau_plan_forward_staged(mau_plan* plan, void* const* state_dev, const void* maps_nhwc_dev, const float* temp_series_dev,
                             const float* metadata_dev, float* out_dev, void* stream) {
  if (!plan || !state_dev || !maps_nhwc_dev || !out_dev) return fail("mau_plan_forward_staged: null argument");
  Ctx c;
  c.state = state_dev; c.maps_staged = maps_nhwc_dev; c.series = temp_series_dev; c.md = metadata_dev; c.out = out_dev;
  c.st = static_cast<cudaStream_t>(stream);
  plan->impl.last_state.assign(state_dev, state_dev + plan->impl.state.size());
  plan->impl.last_series = temp_series_dev; plan->impl.last_md = metadata_dev;
  return plan->impl.run_forward(c);
}

int mau_plan_backward(mau_plan* plan, const float* grad_out_dev, void* const* grads_dev, void* stream) {
  if (!plan || !grad_out_dev || !grads_dev) return fail("mau_plan_backward: null argument");
  Ctx c;
  c.grads = grads_dev; c.gout = grad_out_dev; c.st = static_cast<cudaStream_t>(stream);
  // parameters are read again in backward (dgrad weight pack, BN gamma): reuse the forward's pointers
  c.state = plan->impl.last_state.data();
  c.series = plan->impl.last_series; c.md = plan->impl.last_md;
  return plan->impl.run_backward(c);
}

int mau_plan_buffer_ptr(mau_plan* plan, const char* name, int which, void** ptr_dev, size_t* bytes) {
  if (!plan || !name || !ptr_dev) return fail("mau_plan_buffer_ptr: null argument");
  for (const Buf& b : plan->impl.bufs)
    if (b.name == name) {
      *ptr_dev = which ? b.gptr : b.ptr;
      if (bytes) *bytes = b.bytes;
      return 0;
    }
  return fail("mau_plan_buffer_ptr: no buffer named '%s'", name);
}

int mau_plan_set_state_version(mau_plan* plan, uint64_t version) {
  if (!plan) return fail("null plan");
  plan->impl.state_version = version;
  return 0;
}

int mau_plan_wait_backward_streams(mau_plan* plan, void* stream) {
  if (!plan) return fail("null plan");
  return plan->impl.w_join(static_cast<cudaStream_t>(stream));
}

int mau_plan_set_grad_hook(mau_plan* plan, mau_grad_ready_fn fn, void* user) {
  if (!plan) return fail("null plan");
  plan->impl.hook = fn; plan->impl.hook_user = user;
  return 0;
}

int mau_plan_set_stats_sync(mau_plan* plan, mau_stats_sync_fn fn, void* user, int world_size) {
  if (!plan) return fail("null plan");
  if (fn && world_size < 1) return fail("stats sync: world_size must be >= 1");
  plan->impl.sync_fn = fn; plan->impl.sync_user = user; plan->impl.sync_world = fn ? world_size : 1;
  return 0;
}

int mau_plan_profile(mau_plan* plan, int enable) {
  if (!plan) return fail("null plan");
  plan->impl.profiling = enable != 0;
  plan->impl.prof.clear();
  return 0;
}

int mau_plan_profile_read(mau_plan* plan, char* names, size_t names_len, float* ms, int max_n, int* n) {
  if (!plan || !names || !ms || !n) return fail("null argument");
  std::string s;
  int k = 0;
  for (auto& e : plan->impl.prof) {
    if (k >= max_n || s.size() + e.first.size() + 2 > names_len) break;
    s += e.first; s += "\n";
    ms[k++] = e.second;
  }
  memcpy(names, s.c_str(), s.size() + 1);
  *n = k;
  return 0;
}

int mau_loss_forward_backward(int kind, const float* pred_dev, const float* target_dev, int B, int C, int H, int W,
                              float lambda_grad, float* losses_dev, float* grad_dev, void* stream) {
  if (!pred_dev || !target_dev || !losses_dev) return fail("loss: null argument");
  return op_loss(kind, pred_dev, target_dev, B, C, H, W, lambda_grad, losses_dev, grad_dev,
                 static_cast<cudaStream_t>(stream));
}

int mau_loss_backward(int kind, const float* pred_dev, const float* target_dev, int B, int C, int H, int W,
                      float lambda_total, const float* g_total_dev, const float* g_pixel_dev, const float* g_grad_dev,
                      float* grad_dev, void* stream) {
  if (!pred_dev || !target_dev || !grad_dev) return fail("loss backward: null argument");
  return op_loss_backward(kind, pred_dev, target_dev, B, C, H, W, lambda_total, g_total_dev, g_pixel_dev, g_grad_dev, grad_dev,
                          static_cast<cudaStream_t>(stream));
}

int mau_eval_metrics(const float* maps_dev, int maps_channels, const float* pred_dev, const float* target_dev, int B,
                     int C, int H, int W, float temp_mean, float temp_std, int64_t* dw_map_dev, double* sums_dev,
                     void* stream) {
  if (!maps_dev || !pred_dev || !target_dev || !dw_map_dev || !sums_dev) return fail("eval_metrics: null argument");
  return op_eval_metrics(maps_dev, maps_channels, pred_dev, target_dev, B, C, H, W, temp_mean, temp_std,
                         reinterpret_cast<long long*>(dw_map_dev), sums_dev, static_cast<cudaStream_t>(stream));
}

int mau_laplacian_sums(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, float temp_mean,
                       float temp_std, double* sums_dev, void* stream) {
  if (!pred_dev || !target_dev || !sums_dev) return fail("laplacian_sums: null argument");
  return op_laplacian_sums(pred_dev, target_dev, B, C, H, W, temp_mean, temp_std, sums_dev, static_cast<cudaStream_t>(stream));
}

int64_t mau_ssim_work_floats(int B, int H, int W) { return ssim_work_floats(B, H, W); }

int mau_ssim_loss(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, float* loss_dev, float* grad_dev,
                  float* work_dev, double* acc_dev, void* stream) {
  if (!pred_dev || !target_dev || !loss_dev || !work_dev || !acc_dev) return fail("ssim_loss: null argument");
  return op_ssim_loss(pred_dev, target_dev, B, C, H, W, loss_dev, grad_dev, work_dev, acc_dev, static_cast<cudaStream_t>(stream));
}

int mau_ssim_forward(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, float* loss_dev, float* work_dev,
                     double* acc_dev, void* stream) {
  if (!pred_dev || !target_dev || !loss_dev || !work_dev || !acc_dev) return fail("ssim_forward: null argument");
  return op_ssim_forward(pred_dev, target_dev, B, C, H, W, loss_dev, work_dev, acc_dev, static_cast<cudaStream_t>(stream));
}
int mau_ssim_backward(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, const float* work_dev,
                      const float* upstream_dev, float* grad_dev, void* stream) {
  if (!pred_dev || !target_dev || !work_dev || !grad_dev) return fail("ssim_backward: null argument");
  return op_ssim_backward(pred_dev, target_dev, B, C, H, W, work_dev, upstream_dev, grad_dev, static_cast<cudaStream_t>(stream));
}

int mau_adamw_step(int n_tensors, void* const* params_dev, void* const* grads_dev, void* const* exp_avg_dev,
                   void* const* exp_avg_sq_dev, const int64_t* numels, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, void* stream) {
  if (n_tensors < 0 || (n_tensors && (!params_dev || !grads_dev || !exp_avg_dev || !exp_avg_sq_dev || !numels)))
    return fail("adamw: null argument");
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
  return op_adamw_step(n_tensors, params_dev, grads_dev, exp_avg_dev, exp_avg_sq_dev,
                       reinterpret_cast<const long long*>(numels), lr, beta1, beta2, eps, weight_decay, (long long)step,
                       static_cast<cudaStream_t>(stream));
}

int mau_cast_f32_bf16(const float* src_dev, void* dst_bf16_dev, int64_t n, int max_blocks, void* stream) {
  if (!src_dev || !dst_bf16_dev) return fail("cast: null argument");
  return op_cast_f32_bf16(src_dev, dst_bf16_dev, (long long)n, max_blocks, static_cast<cudaStream_t>(stream));
}
int mau_cast_bf16_f32(const void* src_bf16_dev, float* dst_dev, int64_t n, int max_blocks, void* stream) {
  if (!src_bf16_dev || !dst_dev) return fail("cast: null argument");
  return op_cast_bf16_f32(src_bf16_dev, dst_dev, (long long)n, max_blocks, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- single operators (parity tests)
static View mkview(const void* p, int B, int H, int W, int C, int cs) {
  View v; v.ptr = const_cast<void*>(p); v.B = B; v.H = H; v.W = W; v.C = C; v.cs = cs; v.c0 = 0;
  return v;
}

int mau_op_conv3x3(int impl, int dtype, const void* x_dev, int B, int H, int W, int Cin, int Cin_stride,
                   const float* w_oihw_dev, const float* scale_dev, const float* shift_dev, int relu, int Cout,
                   void* y_dev, int Cout_stride, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Kp = round_up(Cin, 64);
  std::vector<int> kmap(Kp);
  for (int i = 0; i < Kp; ++i) kmap[i] = i < Cin ? i : -1;
  int* kmap_dev = nullptr; void* wp = nullptr;
  MAU_CUDA(cudaMalloc(&kmap_dev, sizeof(int) * Kp));
  MAU_CUDA(cudaMemcpy(kmap_dev, kmap.data(), sizeof(int) * Kp, cudaMemcpyHostToDevice));
  const View x = mkview(x_dev, B, H, W, Cin, Cin_stride), y = mkview(y_dev, B, H, W, Cout, Cout_stride);
  const int zero = 0;
  int rc = 0;
  if (impl == 2) {
    MAU_CUDA(cudaMalloc(&wp, sizeof(float) * 9 * Kp * Cout));
    ConvFfmaParams p;
    rc = conv_ffma_pack_fwd(w_oihw_dev, Cout, Cin, kmap_dev, Kp, static_cast<float*>(wp), st);
    if (!rc) rc = conv_ffma_prepare(&p, x, 1, &zero, &Cin, static_cast<float*>(wp), Kp, Cout, y, scale_dev, shift_dev, relu, 0);
    if (!rc) rc = conv_ffma_launch(dtype, p, B, st);
  } else {
    if (dtype != DT_BF16) { rc = fail("tcgen05 convolution is bf16 only"); }
    else {
      MAU_CUDA(cudaMalloc(&wp, (size_t)2 * 9 * Kp * Cout));
      ConvTcOp op;
      const int mode = (impl & 3) == 0 ? MODE_HALO : ((impl & 3) == 1 ? MODE_TAP : MODE_ROW3);
      rc = conv_tc_pack_fwd(w_oihw_dev, Cout, Cin, kmap_dev, Kp, wp, st);
      if (!rc) rc = conv_tc_prepare(&op, x, 1, &zero, &Cin, wp, Kp, Cout, y, mode, scale_dev, shift_dev, relu, 0, (impl >> 2) & 1);
      if (!rc) rc = conv_tc_launch(op, st);
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(kmap_dev); cudaFree(wp);
  if (!rc && e != cudaSuccess) rc = fail("conv3x3 execution failed: %s", cudaGetErrorString(e));
  return rc;
}

// data gradient of the 3x3 convolution w.r.t. an input-channel range [ci0, ci0 + n_ci): dx = conv(dz, W^T flipped).
// impl 0: tcgen05 halo kernel reading the forward pack MN-major; impl 1: same kernel on a transposed re-pack.
int mau_op_conv3x3_dgrad(int impl, const void* dz_dev, int B, int H, int W, int Cout, int Cout_stride,
                         const float* w_oihw_dev, int Cin, int ci0, int n_ci, void* dx_dev, int dx_stride, int accumulate,
                         void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Kp = round_up(Cin, 64), Kd = round_up(Cout, 64);
  std::vector<int> kmap(Kp);
  for (int i = 0; i < Kp; ++i) kmap[i] = i < Cin ? i : -1;
  int* kmap_dev = nullptr; void* wp = nullptr; void* wd = nullptr;
  MAU_CUDA(cudaMalloc(&kmap_dev, sizeof(int) * Kp));
  MAU_CUDA(cudaMemcpy(kmap_dev, kmap.data(), sizeof(int) * Kp, cudaMemcpyHostToDevice));
  MAU_CUDA(cudaMalloc(&wp, (size_t)2 * 9 * Kp * Cout));
  MAU_CUDA(cudaMalloc(&wd, (size_t)2 * 9 * Kd * n_ci));
  const View dz = mkview(dz_dev, B, H, W, Cout, Cout_stride), dx = mkview(dx_dev, B, H, W, n_ci, dx_stride);
  ConvTcOp op;
  int rc = 0;
  if (impl == 0) {
    rc = conv_tc_pack_fwd(w_oihw_dev, Cout, Cin, kmap_dev, Kp, wp, st);
    if (!rc) rc = conv_tc_prepare_dgrad(&op, dz, wp, Kp, ci0, dx, accumulate);
  } else {
    const int zero = 0;
    rc = conv_tc_pack_dgrad(w_oihw_dev, Cout, Cin, ci0, n_ci, Kd, wd, st);
    if (!rc) rc = conv_tc_prepare(&op, dz, 1, &zero, &Cout, wd, Kd, n_ci, dx, MODE_HALO, nullptr, nullptr, 0, accumulate, 0);
  }
  if (!rc) rc = conv_tc_launch(op, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(kmap_dev); cudaFree(wp); cudaFree(wd);
  if (!rc && e != cudaSuccess) rc = fail("conv3x3 dgrad execution failed: %s", cudaGetErrorString(e));
  return rc;
}

// timing helper for tools/conv_bench.py: prepares once, launches `iters` times between CUDA events
int mau_op_conv3x3_bench(int impl, const void* x_dev, int B, int H, int W, int Cin, int Cin_stride,
                         const float* w_oihw_dev, int Cout, void* y_dev, int Cout_stride, int iters, float* ms_out) {
  const int Kp = round_up(Cin, 64);
  std::vector<int> kmap(Kp);
  for (int i = 0; i < Kp; ++i) kmap[i] = i < Cin ? i : -1;
  int* kmap_dev = nullptr; void* wp = nullptr;
  MAU_CUDA(cudaMalloc(&kmap_dev, sizeof(int) * Kp));
  MAU_CUDA(cudaMemcpy(kmap_dev, kmap.data(), sizeof(int) * Kp, cudaMemcpyHostToDevice));
  MAU_CUDA(cudaMalloc(&wp, (size_t)2 * 9 * Kp * Cout));
  const View x = mkview(x_dev, B, H, W, Cin, Cin_stride), y = mkview(y_dev, B, H, W, Cout, Cout_stride);
  const int zero = 0;
  ConvTcOp op;
  const int mode = (impl & 3) == 0 ? MODE_HALO : ((impl & 3) == 1 ? MODE_TAP : MODE_ROW3);
  int rc = conv_tc_pack_fwd(w_oihw_dev, Cout, Cin, kmap_dev, Kp, wp, 0);
  if (!rc) rc = conv_tc_prepare(&op, x, 1, &zero, &Cin, wp, Kp, Cout, y, mode, nullptr, nullptr, 1, 0, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2 && !rc; ++i) rc = conv_tc_launch(op, 0);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < iters && !rc; ++i) rc = conv_tc_launch(op, 0);
  cudaEventRecord(e1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = ms / (iters > 0 ? iters : 1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(kmap_dev); cudaFree(wp);
  if (!rc && e != cudaSuccess) rc = fail("conv bench failed: %s", cudaGetErrorString(e));
  return rc;
}

int mau_op_conv3x3_wgrad(int impl, int dtype, const void* x_dev, const void* dy_dev, int B, int H, int W, int Cin,
                         int Cin_stride, int Cout, int Cout_stride, float* dw_oihw_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const View x = mkview(x_dev, B, H, W, Cin, Cin_stride), dy = mkview(dy_dev, B, H, W, Cout, Cout_stride);
  MAU_CUDA(cudaMemsetAsync(dw_oihw_dev, 0, sizeof(float) * (size_t)Cout * Cin * 9, st));
  if (impl == 2) return wgrad_ffma_launch(dtype, x, dy, 0, Cin, dw_oihw_dev, 1, st);
  if (dtype != DT_BF16) return fail("tcgen05 wgrad is bf16 only");
  WgradTcOp op;
  if (impl == 1) {
    MAU_TRY(wgrad_tc_prepare(&op, x, dy, 0, Cin, nullptr, 0));
    return wgrad_tc_launch(op, dw_oihw_dev, st);
  }
  const int swap = impl == 4 ? 0 : (impl == 5 ? 1 : (impl == 6 ? 2 : (impl == 7 ? 3 : wgrad_tc_pick_swap(Cout, 1, &Cin))));
  const size_t fl = wgrad_tc_workspace_floats(Cout, Cin, swap);
  float* ws = nullptr;
  MAU_CUDA(cudaMalloc(&ws, sizeof(float) * fl));
  int rc = 0;
  if (cudaMemsetAsync(ws, 0, sizeof(float) * fl, st) != cudaSuccess) rc = fail("wgrad workspace memset failed");
  if (!rc) rc = wgrad_tc_prepare(&op, x, dy, 0, Cin, ws, swap);
  if (!rc) rc = wgrad_tc_launch(op, dw_oihw_dev, st);
  if (!rc) rc = wgrad_tc_finalize(ws, swap, Cout, Cin, dw_oihw_dev, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(ws);
  if (!rc && e != cudaSuccess) rc = fail("wgrad execution failed: %s", cudaGetErrorString(e));
  return rc;
}

// timing helper for tools/conv_bench.py: weight gradient, `iters` launches (memset + kernel + finalize) between events
int mau_op_conv3x3_wgrad_bench(int impl, const void* x_dev, const void* dy_dev, int B, int H, int W, int Cin,
                               int Cin_stride, int Cout, int Cout_stride, float* dw_oihw_dev, int iters, float* ms_out) {
  const View x = mkview(x_dev, B, H, W, Cin, Cin_stride), dy = mkview(dy_dev, B, H, W, Cout, Cout_stride);
  WgradTcOp op;
  const bool v1 = impl == 1;
  const int swap = impl == 4 ? 0 : (impl == 5 ? 1 : (impl == 6 ? 2 : (impl == 7 ? 3 : wgrad_tc_pick_swap(Cout, 1, &Cin))));
  const size_t fl = wgrad_tc_workspace_floats(Cout, Cin, swap);
  float* ws = nullptr;
  if (!v1) MAU_CUDA(cudaMalloc(&ws, sizeof(float) * fl));
  int rc = wgrad_tc_prepare(&op, x, dy, 0, Cin, ws, swap);
  auto once = [&]() -> int {
    if (v1) { if (cudaMemsetAsync(dw_oihw_dev, 0, sizeof(float) * (size_t)Cout * Cin * 9, 0) != cudaSuccess) return fail("memset"); }
    else if (cudaMemsetAsync(ws, 0, sizeof(float) * fl, 0) != cudaSuccess) return fail("memset");
    MAU_TRY(wgrad_tc_launch(op, dw_oihw_dev, 0));
    if (!v1) MAU_TRY(wgrad_tc_finalize(ws, swap, Cout, Cin, dw_oihw_dev, 0));
    return 0;
  };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2 && !rc; ++i) rc = once();
  cudaEventRecord(e0, 0);
  for (int i = 0; i < iters && !rc; ++i) rc = once();
  cudaEventRecord(e1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = ms / (iters > 0 ? iters : 1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(ws);
  if (!rc && e != cudaSuccess) rc = fail("wgrad bench failed: %s", cudaGetErrorString(e));
  return rc;
}

// timing helper (tools/bw_bench.py): one bandwidth-bound kernel of the path, `iters` launches between CUDA
// events.  Every tensor is allocated `sets` times (sets * bytes > L2) and launches rotate over the copies, so
// no launch re-reads data a previous launch left in L2.  Returns ms per launch.
int mau_op_bw_bench(int kind, int dtype, int B, int H, int W, int C, int iters, int sets, float* ms_out) {
  if (sets < 1 || sets > 16 || iters < 1) return fail("bw_bench: bad sets/iters");
  const size_t es = dtype_size(dtype);
  const int H2 = H / 2, W2 = W / 2;
  const size_t big = (size_t)B * H * W * C * es, small = (size_t)B * H2 * W2 * C * es;
  const size_t plane = (size_t)B * H * W * sizeof(float);
  std::vector<void*> allocs;
  auto dalloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, bytes ? bytes : 16);
    allocs.push_back(p);
    return p;
  };
  struct Set { void *a, *b, *c, *s; float *f0, *f1; double* d; };
  std::vector<Set> S(sets);
  for (auto& s : S) {
    s.a = dalloc(big); s.b = dalloc(big); s.c = dalloc(big); s.s = dalloc(small);
    s.f0 = static_cast<float*>(dalloc(std::max(plane * 8, (size_t)B * 23 * H * W * 4)));
    s.f1 = static_cast<float*>(dalloc(plane * 8));
    s.d = static_cast<double*>(dalloc(sizeof(double) * 4 * (C + 64)));
    if (!s.a || !s.b || !s.c || !s.s || !s.f0 || !s.f1 || !s.d) {
      for (void* p : allocs) cudaFree(p);
      return fail("bw_bench: out of device memory");
    }
  }
  std::vector<float> ones(C, 1.f);
  float* coef = static_cast<float*>(dalloc(sizeof(float) * C * 8));
  for (int i = 0; i < 8; ++i) cudaMemcpy(coef + i * C, ones.data(), sizeof(float) * C, cudaMemcpyHostToDevice);
  // bilinear tables (H2,W2) -> (H,W)
  BilinearHost hy, hx;
  bilinear_axis_tables(H2, H, &hy);
  bilinear_axis_tables(W2, W, &hx);
  BilinearTables t;
  t.Hin = H2; t.Win = W2; t.Hout = H; t.Wout = W; t.max_fan_w = hx.max_fan;
  auto up = [&](const void* src, size_t bytes) -> void* { void* d = dalloc(bytes); if (d) cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice); return d; };
  t.ty_off = (int*)up(hy.t_off.data(), 4 * hy.t_off.size()); t.ty_idx = (int*)up(hy.t_idx.data(), 4 * hy.t_idx.size());
  t.ty_w = (float*)up(hy.t_w.data(), 4 * hy.t_w.size());
  t.tx_off = (int*)up(hx.t_off.data(), 4 * hx.t_off.size()); t.tx_idx = (int*)up(hx.t_idx.data(), 4 * hx.t_idx.size());
  t.tx_w = (float*)up(hx.t_w.data(), 4 * hx.t_w.size());
  auto run = [&](const Set& s) -> int {
    const View a = mkview(s.a, B, H, W, C, C), b = mkview(s.b, B, H, W, C, C), c = mkview(s.c, B, H, W, C, C);
    const View sm = mkview(s.s, B, H2, W2, C, C);
    switch (kind) {
      case 0: return op_bn_stats(dtype, a, s.d, 0);
      case 1: return op_bn_apply_relu(dtype, a, coef, coef + C, b, 0);
      case 2: return op_bn_bwd_reduce(dtype, a, b, coef, coef + C, coef + 2 * C, coef + 3 * C, s.d, 0);
      case 3: return op_bn_bwd_apply(dtype, a, b, coef, coef + C, coef + 2 * C, coef + 3 * C, coef + 4 * C, s.d, (long long)B * H * W, b, nullptr, nullptr, nullptr, nullptr, 0);
      case 4: return op_maxpool(dtype, a, sm, 0);
      case 5: return op_maxpool_bwd(dtype, a, sm, nullptr, b, 0);
      case 6: return op_bilinear(dtype, sm, a, t, 0);
      case 7: return op_bilinear_bwd(dtype, a, sm, t, 0, 0);
      case 8: return op_head(dtype, a, coef, coef + 2 * C, 2, 1, s.f0, 0);
      case 9: return op_head_bwd(dtype, a, coef, 2, 1, s.f0, s.f1, b, coef + 5 * C, coef + 7 * C, 0);
      case 10: return op_nchw_to_nhwc(dtype, s.f0, B, 23, H, W, mkview(s.a, B, H, W, 23, 24), 0);
      case 11: return op_embed_broadcast(dtype, coef, 0, a, 0);
      case 12: return op_loss(0, s.f0, s.f1, B, 2, H, W, 0.1f, reinterpret_cast<float*>(s.d), s.f0 + 4 * (size_t)B * H * W, 0);
      case 13: return op_copy_slice(dtype, a, b, 0, 0);     // plain device copy through the same vector path (reference point)
      default: return fail("bw_bench: unknown kind %d", kind);
    }
  };
  int rc = 0;
  for (int i = 0; i < sets && !rc; ++i) rc = run(S[i]);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < iters && !rc; ++i) rc = run(S[i % sets]);
  cudaEventRecord(e1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (void* p : allocs) cudaFree(p);
  if (!rc && e != cudaSuccess) rc = fail("bw_bench failed: %s", cudaGetErrorString(e));
  return rc;
}

int mau_op_maxpool2x2(int dtype, const void* x_dev, int B, int H, int W, int C, void* y_dev, void* stream) {
  return op_maxpool(dtype, mkview(x_dev, B, H, W, C, C), mkview(y_dev, B, H / 2, W / 2, C, C),
                    static_cast<cudaStream_t>(stream));
}

int mau_op_bilinear(int dtype, const void* x_dev, int B, int Hin, int Win, int C, int Hout, int Wout, void* y_dev,
                    void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BilinearHost hy, hx;
  bilinear_axis_tables(Hin, Hout, &hy);
  bilinear_axis_tables(Win, Wout, &hx);
  BilinearTables t;
  t.Hin = Hin; t.Win = Win; t.Hout = Hout; t.Wout = Wout;
  std::vector<void*> tmp;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(bytes, 4)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    tmp.push_back(d);
    return d;
  };
  t.y0 = (int*)up(hy.i0.data(), 4 * hy.i0.size()); t.y1 = (int*)up(hy.i1.data(), 4 * hy.i1.size());
  t.ly = (float*)up(hy.l.data(), 4 * hy.l.size());
  t.x0 = (int*)up(hx.i0.data(), 4 * hx.i0.size()); t.x1 = (int*)up(hx.i1.data(), 4 * hx.i1.size());
  t.lx = (float*)up(hx.l.data(), 4 * hx.l.size());
  int rc = op_bilinear(dtype, mkview(x_dev, B, Hin, Win, C, C), mkview(y_dev, B, Hout, Wout, C, C), t, st);
  cudaStreamSynchronize(st);
  for (void* p : tmp) cudaFree(p);
  return rc;
}

int mau_op_bilinear_bwd(int dtype, const void* gy_dev, int B, int Hin, int Win, int C, int Hout, int Wout, void* gx_dev,
                        int accumulate, int form, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (form < 0 || form > 2) return fail("bilinear_bwd: form must be 0, 1 or 2");
  BilinearHost hy, hx;
  bilinear_axis_tables(Hin, Hout, &hy);
  bilinear_axis_tables(Win, Wout, &hx);
  BilinearTables t;
  t.Hin = Hin; t.Win = Win; t.Hout = Hout; t.Wout = Wout;
  t.max_fan_w = form == 2 ? (1 << 30) : hx.max_fan;        // form 2: the table-driven general kernel
  t.no_stream = form == 1;
  std::vector<void*> tmp;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(bytes, 4)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    tmp.push_back(d);
    return d;
  };
  t.ty_off = (int*)up(hy.t_off.data(), 4 * hy.t_off.size()); t.ty_idx = (int*)up(hy.t_idx.data(), 4 * hy.t_idx.size());
  t.ty_w = (float*)up(hy.t_w.data(), 4 * hy.t_w.size());
  t.tx_off = (int*)up(hx.t_off.data(), 4 * hx.t_off.size()); t.tx_idx = (int*)up(hx.t_idx.data(), 4 * hx.t_idx.size());
  t.tx_w = (float*)up(hx.t_w.data(), 4 * hx.t_w.size());
  int rc = (t.ty_off && t.ty_idx && t.ty_w && t.tx_off && t.tx_idx && t.tx_w) ? 0 : fail("bilinear_bwd: table allocation failed");
  if (!rc) rc = op_bilinear_bwd(dtype, mkview(gy_dev, B, Hout, Wout, C, C), mkview(gx_dev, B, Hin, Win, C, C), t, accumulate, st);
  cudaStreamSynchronize(st);
  for (void* p : tmp) cudaFree(p);
  return rc;
}

int mau_op_nchw_to_nhwc(int dtype, const float* x_dev, int B, int C, int H, int W, int Cstride, void* y_dev,
                        void* stream) {
  return op_nchw_to_nhwc(dtype, x_dev, B, C, H, W, mkview(y_dev, B, H, W, C, Cstride), static_cast<cudaStream_t>(stream));
}
int mau_op_nhwc_to_nchw(int dtype, const void* x_dev, int B, int C, int H, int W, int Cstride, float* y_dev,
                        void* stream) {
  return op_nhwc_to_nchw(dtype, mkview(x_dev, B, H, W, C, Cstride), y_dev, static_cast<cudaStream_t>(stream));
}
int mau_op_lstm_last_hidden(const float* series_dev, int B, int T, int hidden, const float* w_ih, const float* w_hh,
                            const float* b_ih, const float* b_hh, float* h_out_dev, void* stream) {
  return op_lstm_fwd(series_dev, B, T, hidden, w_ih, w_hh, b_ih, b_hh, h_out_dev, nullptr,
                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
