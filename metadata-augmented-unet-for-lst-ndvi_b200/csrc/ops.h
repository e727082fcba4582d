// ops.h -- host launchers of the bandwidth-bound kernels (elementwise.cu, norm.cu, encoders.cu,
// loss.cu, metrics.cu).  Every function enqueues on `st` and returns 0 / non-zero.
#pragma once
#include <vector>
#include "common.h"
#include "bilinear_tables.h"

namespace mau {

// ---- layout (elementwise.cu) ----------------------------------------------------------------
int op_nchw_to_nhwc(int dt, const float* x, int B, int C, int H, int W, const View& y, cudaStream_t st);
int op_nhwc_to_nchw(int dt, const View& x, float* y, cudaStream_t st);

// ---- max-pool 2x2/2 floor mode (nn.MaxPool2d(2,2), reference src/model.py:218/58) -------------
int op_maxpool(int dt, const View& x, const View& y, cudaStream_t st);
// gx = (addend ? addend : 0) + route(gy) ; tie -> first max in row-major window order
int op_maxpool_bwd(int dt, const View& x, const View& gy, const View* addend, const View& gx, cudaStream_t st);

// ---- bilinear align_corners=True (reference src/model.py:121,219,245) --------------------------
int op_bilinear(int dt, const View& x, const View& y, const BilinearTables& t, cudaStream_t st);
int op_bilinear_bwd(int dt, const View& gy, const View& gx, const BilinearTables& t, int accumulate,
                    cudaStream_t st);

// ---- embedding broadcast / reduction (reference src/model.py:248-259, 98-108) -----------------
// y[b,h,w,c] = emb[b*emb_stride + c]
int op_embed_broadcast(int dt, const float* emb, int emb_stride, const View& y, cudaStream_t st);
// demb[b*stride + c] (+)= sum_{h,w} g[b,h,w,c]
int op_embed_reduce(int dt, const View& g, float* demb, int emb_stride, int accumulate, cudaStream_t st);

// ---- dst[b] = src[0] for every batch row b of dst (shared-maps sweep: one-tile tensor -> concat slice) ----
int op_broadcast_batch(int dt, const View& src, const View& dst, cudaStream_t st);

// ---- dst (=|+=) src on channel-slice views ------------------------------------------------------
int op_copy_slice(int dt, const View& src, const View& dst, int accumulate, cudaStream_t st);

// ---- 1x1 head + tanh on channel 0 (reference src/model.py:284-292) -----------------------------
int op_head(int dt, const View& x, const float* w, const float* bias, int OC, int apply_tanh, float* out_nchw,
            cudaStream_t st);
// gx = W^T (gout * act'), dW += ..., db += ... (dw/db must be zeroed by the caller)
int op_head_bwd(int dt, const View& x, const float* w, int OC, int apply_tanh, const float* out_nchw,
                const float* gout_nchw, const View& gx, float* dw, float* db, cudaStream_t st);

// ---- BatchNorm (norm.cu; nn.BatchNorm2d of reference src/model.py:13,15) ------------------------
// eval: scale = gamma / sqrt(rv + eps), shift = (conv_bias - rm) * scale + beta
int op_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                    const float* conv_bias, int C, float eps, float* scale, float* shift, cudaStream_t st);
// train: per-channel sum / sum of squares of z into sums[2*C] (double, zeroed by the caller)
int op_bn_stats(int dt, const View& z, double* sums, cudaStream_t st);
// mean/var -> scale, shift (for the apply pass), saved mean / rstd, running-stat update
int op_bn_finalize_train(const double* sums, long long count, const float* gamma, const float* beta, int C,
                         float eps, float momentum, float* running_mean, float* running_var, float* scale,
                         float* shift, float* save_mean, float* save_rstd, cudaStream_t st);
// op_bn_finalize_train + op_bn_apply_relu in one launch (every block derives the coefficients of its channels)
int op_bn_finalize_apply_relu(int dt, const View& z, const double* sums, long long count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                              float* scale, float* shift, float* save_mean, float* save_rstd, const View& y,
                              cudaStream_t st);
// y = relu(z * scale + shift) into a (slice) view
int op_bn_apply_relu(int dt, const View& z, const float* scale, const float* shift, const View& y,
                     cudaStream_t st);
// backward, pass 1: sums[0..C) = sum g~, sums[C..2C) = sum g~ * xhat, g~ = gy * (y > 0); the ReLU mask is
// recomputed from z with the forward's scale / shift, so y is not read
int op_bn_bwd_reduce(int dt, const View& gy, const View& z, const float* scale, const float* shift, const float* mean,
                     const float* rstd, double* sums, cudaStream_t st);
// backward, pass 2: dz = gamma*rstd*(g~ - s1/n - xhat*s2/n) written over dz_out (may alias z).  sum(dz), the
// gradient of the convolution bias, is identically zero under batch statistics and is not reduced.
// block 0 also writes dgamma = param_sums[C..2C), dbeta = param_sums[0..C), dbias = 0 when param_sums != nullptr
int op_bn_bwd_apply(int dt, const View& gy, const View& z, const float* scale, const float* shift, const float* gamma,
                    const float* mean, const float* rstd, const double* sums, long long count, const View& dz_out,
                    const double* param_sums, float* dgamma, float* dbeta, float* dbias, cudaStream_t st);
// dgamma = s2, dbeta = s1, dbias = dbias_sums (nullptr -> 0)
int op_bn_bwd_finalize(const double* sums, const double* dbias_sums, int C, float* dgamma, float* dbeta,
                       float* dbias, cudaStream_t st);
int op_bump_counters(long long* const* counters_dev, int n, cudaStream_t st);

// ---- encoders (encoders.cu) ---------------------------------------------------------------------
// MetadataEncoder: Linear(F,32) -> ReLU -> Linear(32,D)  (reference src/model.py:38-48); fp32
int op_mlp_fwd(const float* md, int B, int F, const float* w0, const float* b0, const float* w2, const float* b2,
               int D, float* hidden /*[B,32]*/, float* out, int out_stride, cudaStream_t st);
int op_mlp_bwd(const float* md, int B, int F, const float* w0, const float* w2, int D, const float* hidden,
               const float* gout, int gout_stride, float* dw0, float* db0, float* dw2, float* db2,
               cudaStream_t st);
// TemporalEncoder: LSTM(1,Hd) last hidden -> Linear(Hd,D)  (reference src/model.py:23-34); fp32
// gates_save (nullable): [B,T,4*Hd] post-activation gates + [B,T,Hd] cell states for BPTT
int op_lstm_fwd(const float* series, int B, int T, int Hd, const float* w_ih, const float* w_hh, const float* b_ih,
                const float* b_hh, float* h_last, float* save /*nullable*/, cudaStream_t st);
int op_lstm_bwd(const float* series, int B, int T, int Hd, const float* w_hh, const float* save,
                const float* dh_last, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, float* scratch,
                cudaStream_t st);
size_t lstm_save_floats(int B, int T, int Hd);
size_t lstm_bwd_scratch_floats(int B, int Hd);
int op_linear_fwd(const float* x, int B, int K, const float* w, const float* b, int N, float* y, int y_stride,
                  cudaStream_t st);
int op_linear_bwd(const float* x, int B, int K, const float* w, int N, const float* gy, int gy_stride, float* gx,
                  float* dw, float* db, cudaStream_t st);

// ---- loss / metrics -----------------------------------------------------------------------------
int op_loss(int kind, const float* pred, const float* tgt, int B, int C, int H, int W, float lambda_grad,
            float* losses, float* grad, cudaStream_t st);
// d(g_total*total + g_pixel*pixel + g_grad*gradient)/d pred with total = pixel + lambda_total*gradient; the g_* are
// nullable DEVICE scalars (autograd's upstream gradients), so no host synchronisation and no separate scaling pass
int op_loss_backward(int kind, const float* pred, const float* tgt, int B, int C, int H, int W, float lambda_total,
                     const float* g_total, const float* g_pixel, const float* g_grad, float* grad, cudaStream_t st);
int op_eval_metrics(const float* maps, int maps_c, const float* pred, const float* tgt, int B, int C, int H, int W,
                    float temp_mean, float temp_std, long long* dw_map, double* sums, cudaStream_t st);
// per (sample, channel): {sum, sum of squares} of scipy.ndimage.laplace for pred and tgt -> out [B*C][4] (test/evaluate.py:241-242)
int op_laplacian_sums(const float* pred, const float* tgt, int B, int C, int H, int W, float temp_mean, float temp_std,
                      double* out, cudaStream_t st);
// SSIM term of the training loss (src/utils/losses.py:72-95): loss[0] = 1 - mean SSIM, grad = d loss / d pred (or null);
// work holds ssim_work_floats(B, H, W) floats, acc one double
long long ssim_work_floats(int B, int H, int W);
void ssim_debug_force_pool(int f);     // tests: force piq's average-pool factor (0 = piq's rule, f = round(min(H,W)/256))
int op_ssim_loss(const float* pred, const float* tgt, int B, int C, int H, int W, float* loss, float* grad, float* work,
                 double* acc, cudaStream_t st);
// the two halves separately (autograd): forward leaves the derivative maps in work; backward multiplies the nullable
// DEVICE scalar `upstream` (d L / d ssim_loss) into the gradient
int op_ssim_forward(const float* pred, const float* tgt, int B, int C, int H, int W, float* loss, float* work, double* acc,
                    cudaStream_t st);
int op_ssim_backward(const float* pred, const float* tgt, int B, int C, int H, int W, const float* work, const float* upstream,
                     float* grad, cudaStream_t st);

// ---- backward of a conv w.r.t. a spatially constant input segment (embgrad.cu; U-Net++ embedding planes) ------
// dz: the conv's output gradient [B,H,W,Cout]; emb [B, emb_stride] holds the segment's E values per image at emb[b*stride + c].
// dw_oihw[:, ci0:ci0+E, :, :] is overwritten, demb[b*stride + c] is accumulated.  scratch: emb_grad_scratch_floats(B, Cout).
size_t emb_grad_scratch_floats(int B, int Cout);
int op_emb_segment_grad(int dt, const View& dz, const float* w_oihw, int Cin, int ci0, int E, const float* emb,
                        int emb_stride, float* dw_oihw, float* demb, float* scratch, cudaStream_t st);

// ---- optimizer (optim.cu): torch.optim.AdamW step for every tensor in one launch ---------------------
int op_adamw_step(int n_tensors, void* const* params, void* const* grads, void* const* exp_avg,
                  void* const* exp_avg_sq, const long long* numels, double lr, double beta1, double beta2, double eps,
                  double weight_decay, long long step, cudaStream_t st);

// fp32 gradient bucket <-> bf16 wire buffer of the data-parallel all-reduce (max_blocks bounds the SMs the cast may take
// next to the backward kernels it overlaps; 0 = default)
int op_cast_f32_bf16(const float* src, void* dst_bf16, long long n, int max_blocks, cudaStream_t st);
int op_cast_bf16_f32(const void* src_bf16, float* dst, long long n, int max_blocks, cudaStream_t st);

}  // namespace mau
