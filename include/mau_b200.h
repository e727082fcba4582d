/*
 * mau_b200.h -- C ABI of the B200-native hot path of the Metadata-Augmented U-Net.
 *
 * This is the drop-in boundary for the forward/backward of the reference's
 * `UrbanPredictor` (reference src/model.py:295-329).  The reference has no FFI of its
 * own (it is pure Python on top of PyTorch/ATen), so each entry point below cites the
 * reference *call* it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; mau_last_error() returns
 *     a thread-local, NUL-terminated description of the last failure.  Nothing throws
 *     across the ABI.
 *   - all pointers named *_dev are device pointers on the plan's device; the caller owns
 *     parameters, buffers, inputs, outputs and gradients; the plan owns only its
 *     workspace, packed weights and TMA descriptors.
 *   - `stream` is a cudaStream_t passed as void*; all work is ordered on it and nothing is
 *     synchronised with the host (the plan is stateful: one forward/backward in flight per plan).
 *     A plan may run its LSTM / MLP encoder kernels on a private side stream; they are forked from
 *     and joined back into `stream` with events inside the same call.
 *   - tensors crossing the ABI are contiguous fp32 NCHW exactly as the reference's
 *     callers hold them (src/dataset.py:99-106); counters are int64.
 */
#ifndef MAU_B200_H_
#define MAU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAU_MODEL_UNET   0   /* UrbanPredictor_unet,   src/model.py:195-292 */
#define MAU_MODEL_UNETPP 1   /* UrbanPredictor_unetpp, src/model.py:51-193  */

#define MAU_PRECISION_BF16 0 /* NHWC bf16 activations, tcgen05 implicit GEMM, fp32 accumulate */
#define MAU_PRECISION_FP32 1 /* NHWC fp32 activations, FFMA convolutions (1e-5 parity mode)   */

typedef struct mau_config {
  int32_t model_type;           /* MAU_MODEL_* (src/model.py:299,311)                         */
  int32_t spatial_channels;     /* ctor arg, 23 in conf/config.yaml:18                         */
  int32_t temporal_dim;         /* TemporalEncoder out_dim, src/model.py:27                    */
  int32_t meta_features;        /* MetadataEncoder in_features (4 or 8), src/model.py:42       */
  int32_t meta_dim;             /* MetadataEncoder out_dim, src/model.py:44                    */
  int32_t lstm_dim;             /* LSTM hidden size, src/model.py:26                           */
  int32_t out_channels;         /* 1x1 head, src/model.py:241 / :96                            */
  int32_t filters[5];           /* nb_filter, src/model.py:322 / :54                           */
  int32_t temporal_embeddings;  /* U-Net only flag, src/model.py:198; U-Net++ always 1         */
  int32_t metadata_embeddings;  /* U-Net only flag, src/model.py:199; U-Net++ always 1         */
  int32_t deep_supervision;     /* U-Net++ only, src/model.py:52,180                           */
  int32_t batch, height, width; /* maps [B, spatial_channels, H, W]                            */
  int32_t seq_len;              /* T of temp_series [B, T] (runtime length, src/model.py:29)   */
  int32_t training;             /* nn.Module.training: batch-stat BN + saved activations        */
  int32_t precision;            /* MAU_PRECISION_*                                             */
  int32_t device;               /* CUDA device ordinal                                         */
  int32_t flags;                /* MAU_FLAG_* bit set                                          */
} mau_config;

#define MAU_FLAG_SHARED_MAPS   1  /* all B rows share maps and series (metadata_sensitivity sweep,
                                     test/metadata_sensitivity.py:294-311): encoder computed once */
#define MAU_FLAG_CONV_TAPLOAD  2  /* debug: force the 9-box-loads-per-chunk conv main loop        */
#define MAU_FLAG_CONV_FFMA     4  /* debug: run bf16 plans on the FFMA convolution kernels         */
#define MAU_FLAG_HALO_BASEOFF  256 /* debug: halo main loop fills the UMMA descriptor base_offset   */
#define MAU_FLAG_CONV_ROW3     512 /* debug: three-row-box conv main loop instead of the halo kernel */
#define MAU_FLAG_WGRAD_V1      1024 /* debug: first-generation weight-gradient kernel (fp32 atomics)   */
#define MAU_FLAG_NO_CONV_STATS 16384 /* A-B: BatchNorm statistics by a separate pass over z instead of inside the convolution */
#define MAU_FLAG_NO_WGRAD_OVERLAP 8192 /* A-B: weight-gradient kernels on the caller's stream (no second stream) */
#define MAU_FLAG_EMB_DENSE_BWD 2048 /* debug: U-Net++ embedding planes back-propagated densely (dgrad + wgrad launches) */

typedef struct mau_plan mau_plan; /* opaque */

/* library / error plumbing */
const char* mau_last_error(void);
int         mau_version(void);
/* number of CUDA kernels this library has launched in the calling process (bench gpu_launches) */
int64_t     mau_launch_count(void);

/* --- the plan: replaces nn.Module construction + .to(device) (src/train.py:194-206) ------- */
int    mau_plan_create(const mau_config* cfg, mau_plan** out);
int    mau_plan_destroy(mau_plan* plan);
size_t mau_plan_workspace_bytes(const mau_plan* plan);
/* number of state tensors (state_dict order: parameters and buffers) the plan expects */
int    mau_plan_num_state(const mau_plan* plan);
/* element count of state tensor i, and its role: 0 = trainable fp32 parameter used by this
 * configuration, 1 = fp32 parameter present but unused (flag-disabled encoder: grad stays
 * None, src/model.py:263-264), 2 = fp32 running statistic, 3 = int64 num_batches_tracked */
int    mau_plan_state_info(const mau_plan* plan, int i, int64_t* numel, int* role);
/* a human-readable description of the layer graph (JSON), for tests and tooling; works
 * without a GPU when created through mau_plan_describe_config */
int    mau_plan_describe_config(const mau_config* cfg, char* buf, size_t buflen);
/* algorithmic work of one forward over the whole batch (dense reference graph) */
int    mau_plan_flops(const mau_plan* plan, double* fwd_flops, double* bwd_flops);
/* FLOPs the 3x3 convolution kernels of ONE forward / ONE backward actually execute: the dense conv FLOPs unless
 * MAU_FLAG_SHARED_MAPS runs the encoder once for the whole batch (forward) or the U-Net++ embedding planes are
 * back-propagated in closed form (backward; 0 for eval plans) */
int    mau_plan_exec_flops(const mau_plan* plan, double* conv_flops_per_forward, double* conv_flops_per_backward);

/* --- forward: replaces `model(maps, temp_series, metadata)` (src/model.py:328,
 *     called at src/train.py:245, test/evaluate.py:186, test/metadata_sensitivity.py:310) ---
 * state_dev[i]: device pointer of state tensor i (state_dict order).  In training mode the
 * running_mean / running_var / num_batches_tracked entries are updated in place exactly like
 * nn.BatchNorm2d does.  out_dev: [B, out_channels, H, W] fp32 (deep supervision:
 * [4, B, out_channels, H, W]). */
int mau_plan_forward(mau_plan* plan, void* const* state_dev, const float* maps_dev,
                     const float* temp_series_dev, const float* metadata_dev, float* out_dev,
                     void* stream);

/* The same forward with the tiles already STAGED: maps_nhwc_dev is [B, H, W, round_up(spatial_channels, 8)] in the plan's
 * activation type (bf16 for MAU_PRECISION_BF16, fp32 otherwise) -- what a loader produces when it converts while it decodes.
 * Halves the host-to-device bytes of the reference's fp32 NCHW contract (48 MB instead of 92 MB per 16 tiles) and replaces
 * the layout kernel by one device copy; values are identical to mau_plan_forward on the fp32 tiles rounded to nearest-even.
 * (B = 1 for a MAU_FLAG_SHARED_MAPS plan.) */
int mau_plan_forward_staged(mau_plan* plan, void* const* state_dev, const void* maps_nhwc_dev,
                            const float* temp_series_dev, const float* metadata_dev, float* out_dev, void* stream);

/* test / tooling access to the plan's own activation storage: device pointer of the NHWC buffer `name` (names, extents
 * and channel strides are listed under "buffers" by mau_plan_describe_config; bf16 or fp32 elements according to the
 * plan's precision).  which = 0: the activation, 1: its gradient twin (training plans; NULL until the plan has run a
 * backward pass).  Used by the teacher-forced per-layer parity test (tests/test_bf16_layers_gpu.py). */
int mau_plan_buffer_ptr(mau_plan* plan, const char* name, int which, void** ptr_dev, size_t* bytes);

/* optional eval-mode hint: a number that changes whenever any state tensor is modified (e.g. the sum
 * of torch's tensor._version counters).  While it -- and every state pointer -- is unchanged between
 * forwards, the packed bf16 weights and folded BatchNorm vectors are reused; 0 = always re-pack. */
int mau_plan_set_state_version(mau_plan* plan, uint64_t version);

/* --- backward: replaces `loss.backward()` through the model (src/train.py:252) -------------
 * grad_out_dev: dL/d out, same shape as out.  grads_dev[i]: where to write dL/d state[i]
 * (fp32, same shape), or NULL to skip; entries for unused / non-parameter state must be NULL.
 * Gradients are written (not accumulated). Must follow a training-mode forward on the plan. */
int mau_plan_backward(mau_plan* plan, const float* grad_out_dev, void* const* grads_dev,
                      void* stream);

/* optional: called from inside backward as soon as all gradients up to (and including) the
 * given state index are final on `stream` -- the hook a data-parallel caller uses to launch a
 * bucketed all-reduce that overlaps the rest of backward (new capability; the reference is
 * single-GPU, src/train.py:99).  first_index..last_index is a contiguous state-index range. */
typedef void (*mau_grad_ready_fn)(void* user, int first_index, int last_index);
int mau_plan_set_grad_hook(mau_plan* plan, mau_grad_ready_fn fn, void* user);
/* backward runs the weight-gradient kernels on a plan-owned second stream (they overlap the next layer's bandwidth-bound
 * BatchNorm / pool / bilinear backward on `stream`); mau_plan_backward joins it before returning control of `stream`.
 * A grad hook that launches work on ANOTHER stream (the all-reduce) calls this first: `stream` then waits for every
 * weight-gradient launch enqueued so far, in addition to the event it records on the backward stream itself. */
int mau_plan_wait_backward_streams(mau_plan* plan, void* stream);

/* optional SyncBN (new capability): when set, every training-mode BatchNorm calls fn with its
 * device buffer of per-channel sums (double[n], n = 2*C: forward {sum z, sum z^2}, backward
 * {sum g, sum g*xhat}) right after the local reduction; fn must all-reduce (SUM) the buffer in place
 * across the world_size ranks, ordered on the stream the plan call was given.  Batch statistics, running
 * statistics and dz then equal those of one process running the global batch (the reference semantics,
 * src/model.py:13 on one device); dgamma / dbeta stay per-rank so that the gradient average is exact. */
typedef void (*mau_stats_sync_fn)(void* user, void* sums_dev, int n_doubles);
int mau_plan_set_stats_sync(mau_plan* plan, mau_stats_sync_fn fn, void* user, int world_size);

/* Environment switches read by the library (all optional; measurement / A-B knobs, none changes results beyond rounding order):
 *   MAU_SM_RESERVE=n        default of mau_set_sm_reserve (below)
 *   MAU_WHOLE_WAVES=0       equal-work element-wise kernels (BatchNorm apply passes, head backward) keep the grids they
 *                           had before they were rounded down to whole waves of resident blocks
 *   MAU_WGRAD_PAIR=0        weight gradients of the <= 64-channel sides on the split-K kernel instead of the tap-pair kernel
 *   MAU_WGRAD_SWAP=0..3     force the weight-gradient form (0 / 1: split-K kernel, operand orientation; 2 / 3: tap-pair kernel
 *                           with dY / X carrying the tap shift)
 *   MAU_WGRAD_FORK_LATE=1   second (weight-gradient) stream forks behind the data gradient instead of behind the BatchNorm
 *                           backward
 *   MAU_NO_COL3=1, MAU_NO_BRES=1, MAU_CONV_CFG=bn,mt,nbuf   convolution tile selection (column-folded N = 192 form, resident
 *                           weights, explicit tile shape)
 * The Python layer adds MAU_PRECISION (bf16 | fp32) and MAU_FLAGS (MAU_FLAG_* bits or-ed into every plan). */

/* SMs the persistent kernels leave free (default 0, or $MAU_SM_RESERVE): set it to the number of CTAs
 * a concurrently running collective (NCCL all-reduce overlapped with backward) occupies, so that a
 * persistent grid never spills into a second wave.  Applies to the BACKWARD launches (dgrad, wgrad) of
 * plans created afterwards; forward launches always use every SM (no collective runs during forward). */
int mau_set_sm_reserve(int n_sms);

/* per-layer device timings of the last forward/backward (ms, CUDA events on `stream`);
 * enable, run, then read.  names: '\n'-separated, same order as ms[]. */
int mau_plan_profile(mau_plan* plan, int enable);
int mau_plan_profile_read(mau_plan* plan, char* names, size_t names_len, float* ms, int max_n, int* n);

/* --- training loss terms: replaces F.l1_loss / F.mse_loss / gradient_loss
 *     (src/utils/losses.py:5-25,33,67-70); the SSIM term is mau_ssim_loss below.
 * kind: 0 = L1 + lambda*grad, 1 = MSE + lambda*grad.  losses_dev[4] = {total, pixel, gradient, 0}.
 * grad_dev (nullable): dL_total/d pred, same shape as pred. */
int mau_loss_forward_backward(int kind, const float* pred_dev, const float* target_dev, int B, int C,
                              int H, int W, float lambda_grad, float* losses_dev, float* grad_dev,
                              void* stream);

/* backward of the same terms for autograd (src/train.py:252 `batch_loss.backward()` reaches the criterion first):
 * grad_dev = d(g_total*total + g_pixel*pixel + g_gradient*gradient)/d pred with total = pixel + lambda_total*gradient.
 * The g_*_dev are nullable DEVICE scalars (the upstream gradients of the three dictionary entries), read by the kernel:
 * no host synchronisation, no separate scaling pass.  All three entries of the reference's dictionaries stay
 * differentiable this way (src/utils/losses.py:5-25 `gradient_loss` included). */
int mau_loss_backward(int kind, const float* pred_dev, const float* target_dev, int B, int C, int H, int W,
                      float lambda_total, const float* g_total_dev, const float* g_pixel_dev,
                      const float* g_gradient_dev, float* grad_dev, void* stream);

/* --- optimizer step: replaces torch.optim.AdamW(...).step() (src/train.py:213-214,255) -------
 * One launch updates all n_tensors parameter tensors (fp32, contiguous) in place with decoupled weight
 * decay; exp_avg / exp_avg_sq are the optimizer's state tensors (same layout as torch's, so
 * optimizer.state_dict() stays interchangeable); step is the 1-based update count used for the
 * bias corrections.  New capability on the path's "next" row. */
int mau_adamw_step(int n_tensors, void* const* params_dev, void* const* grads_dev, void* const* exp_avg_dev,
                   void* const* exp_avg_sq_dev, const int64_t* numels, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int64_t step, void* stream);

/* fp32 gradient bucket <-> bf16 wire buffer of the data-parallel gradient all-reduce (new capability):
 * halves the NVLink payload; max_blocks bounds the CTAs of the cast next to the backward kernels it overlaps (0 = default) */
int mau_cast_f32_bf16(const float* src_dev, void* dst_bf16_dev, int64_t n, int max_blocks, void* stream);
int mau_cast_bf16_f32(const void* src_bf16_dev, float* dst_dev, int64_t n, int max_blocks, void* stream);

/* --- evaluation metrics: replaces the NumPy loop of test/evaluate.py:210-275 ---------------
 * dw_map_dev [B,H,W] int64 = argmax_c(maps[b,c]*c, c<9) (ties -> lowest index, bit-exact);
 * temperature channel (index 1) is un-normalised by *temp_std + temp_mean (test/evaluate.py:33-36)
 * unless temp_std == 0.  sums_dev [B, C, 10, 3] float64: for class slot k (0 = overall, 1..9 = DW
 * class k-1): {count, sum |p-g|, sum (p-g)^2}. */
int mau_eval_metrics(const float* maps_dev, int maps_channels, const float* pred_dev,
                     const float* target_dev, int B, int C, int H, int W, float temp_mean,
                     float temp_std, int64_t* dw_map_dev, double* sums_dev, void* stream);

/* SSIM term of the training loss, src/utils/losses.py:72-95: channel 0 scaled (x+1)/2, channel 1 clamped to [0,1],
 * piq.ssim(..., data_range=1, reduction='none') -- 11x11 Gaussian window, sigma 1.5, valid filtering, k1 .01, k2 .03 --
 * loss_dev[0] = 1 - mean SSIM over images, channels 0..1 and window positions; grad_dev (may be NULL) [B,C,H,W] =
 * d loss / d pred (zero for channels >= 2).  work_dev: mau_ssim_work_floats(B,H,W) floats of scratch, acc_dev: one double.
 * piq is a third-party dependency the reference does not pin: parity for this term is UNPINNED (see oracle/ssim_oracle.py).
 * Tiles with min(H,W) >= 384 are average-pooled by f = round(min(H,W)/256) first, like piq.ssim(downsample=True) does
 * (the app's 512 x 512 tiles: f = 2). */
int64_t mau_ssim_work_floats(int B, int H, int W);
int mau_ssim_loss(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, float* loss_dev,
                  float* grad_dev, float* work_dev, double* acc_dev, void* stream);

/* The same in two calls for autograd: mau_ssim_forward leaves the per-window derivative maps in work_dev (keep it until the
 * backward); mau_ssim_backward writes grad_dev = upstream * d loss / d pred, upstream_dev being a nullable DEVICE scalar
 * (the gradient autograd hands to the term, e.g. lambda_ssim = 0.5, src/utils/losses.py:95) -- no host synchronisation. */
int mau_ssim_forward(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, float* loss_dev,
                     float* work_dev, double* acc_dev, void* stream);
int mau_ssim_backward(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, const float* work_dev,
                      const float* upstream_dev, float* grad_dev, void* stream);

/* Sharpness metric of the same evaluation loop (test/evaluate.py:241-242): np.var(scipy.ndimage.laplace(x)) of the
 * un-normalised prediction and target of every (sample, channel).  sums_dev [B, C, 4] float64 =
 * {sum L(pred), sum L(pred)^2, sum L(target), sum L(target)^2} with L = the 5-point Laplacian in scipy's default
 * 'reflect' boundary mode, evaluated per axis in double and rounded to fp32 like scipy does;
 * variance = sum2 / (H*W) - (sum / (H*W))^2 is formed by the caller. */
int mau_laplacian_sums(const float* pred_dev, const float* target_dev, int B, int C, int H, int W, float temp_mean,
                       float temp_std, double* sums_dev, void* stream);

/* --- single operators on raw NHWC device buffers (used by the kernel-level parity tests) ----
 * dtype: 0 = bf16, 1 = fp32.  x [B,H,W,Cin_stride], w OIHW fp32 [Cout,Cin,3,3],
 * y [B,H,W,Cout_stride]; y = relu?(conv(x,w)*scale + shift).  impl: 0 = tcgen05 persistent halo
 * kernel (default), 1 = tcgen05 tap-load main loop, 3 = tcgen05 three-row-box main loop, 2 = FFMA. */
int mau_op_conv3x3(int impl, int dtype, const void* x_dev, int B, int H, int W, int Cin, int Cin_stride,
                   const float* w_oihw_dev, const float* scale_dev, const float* shift_dev, int relu,
                   int Cout, void* y_dev, int Cout_stride, void* stream);
/* dX[:, ci0 : ci0+n_ci] (=|+=) data gradient of the same convolution from dZ [B,H,W,Cout_stride] (bf16 NHWC);
 * impl 0 = reads W^T MN-major out of the forward weight pack, 1 = transposed re-pack (first generation) */
int mau_op_conv3x3_dgrad(int impl, const void* dz_dev, int B, int H, int W, int Cout, int Cout_stride,
                         const float* w_oihw_dev, int Cin, int ci0, int n_ci, void* dx_dev, int dx_stride,
                         int accumulate, void* stream);
/* timing helper (tools/conv_bench.py): bf16 tcgen05 conv, `iters` launches between CUDA events */
int mau_op_conv3x3_bench(int impl, const void* x_dev, int B, int H, int W, int Cin, int Cin_stride,
                         const float* w_oihw_dev, int Cout, void* y_dev, int Cout_stride, int iters, float* ms_out);
/* dW [Cout,Cin,3,3] fp32 = sum_pixels dy (x) x ; impl 0 = tcgen05 persistent split-K kernel (orientation
 * or tap-pair variant chosen by the cost model), 4 / 5 = split-K kernel with M = output / input channels forced,
 * 6 / 7 = tap-pair kernel with the output / input side (<= 64 channels) carrying the tap shift,
 * 1 = first-generation tcgen05 kernel (atomics), 2 = FFMA */
int mau_op_conv3x3_wgrad(int impl, int dtype, const void* x_dev, const void* dy_dev, int B, int H, int W,
                         int Cin, int Cin_stride, int Cout, int Cout_stride, float* dw_oihw_dev,
                         void* stream);
/* timing helper (tools/conv_bench.py): weight gradient incl. workspace memset + transpose */
int mau_op_conv3x3_wgrad_bench(int impl, const void* x_dev, const void* dy_dev, int B, int H, int W, int Cin,
                               int Cin_stride, int Cout, int Cout_stride, float* dw_oihw_dev, int iters,
                               float* ms_out);
/* timing helper (tools/bw_bench.py): one bandwidth-bound kernel, rotating over `sets` copies of its tensors
 * (sets * bytes > L2).  kind: 0 bn_stats, 1 bn_apply_relu, 2 bn_bwd_reduce, 3 bn_bwd_apply, 4 maxpool,
 * 5 maxpool_bwd, 6 bilinear (H/2 -> H), 7 bilinear_bwd, 8 head, 9 head_bwd, 10 nchw_to_nhwc (23 ch),
 * 11 embed_broadcast, 12 loss (L1 + gradient, fwd+bwd), 13 slice copy. */
int mau_op_bw_bench(int kind, int dtype, int B, int H, int W, int C, int iters, int sets, float* ms_out);
int mau_op_maxpool2x2(int dtype, const void* x_dev, int B, int H, int W, int C, void* y_dev, void* stream);
int mau_op_bilinear(int dtype, const void* x_dev, int B, int Hin, int Win, int C, int Hout, int Wout,
                    void* y_dev, void* stream);
/* backward of mau_op_bilinear: gx [B,Hin,Win,C] (=|+=) the transposed resize of gy [B,Hout,Wout,C] (autograd of
 * F.interpolate(..., mode="bilinear", align_corners=True), reference src/model.py:12-17).  form: 0 = the kernel the plan
 * picks for this shape (the streaming kernel when rows are up-sampled), 1 = the batched per-pixel gather kernel (what
 * the plan uses when rows are down-sampled and no source column receives more than six contributions), 2 = the
 * table-driven general kernel. */
int mau_op_bilinear_bwd(int dtype, const void* gy_dev, int B, int Hin, int Win, int C, int Hout, int Wout,
                        void* gx_dev, int accumulate, int form, void* stream);
int mau_op_nchw_to_nhwc(int dtype, const float* x_dev, int B, int C, int H, int W, int Cstride,
                        void* y_dev, void* stream);
int mau_op_nhwc_to_nchw(int dtype, const void* x_dev, int B, int C, int H, int W, int Cstride,
                        float* y_dev, void* stream);
int mau_op_lstm_last_hidden(const float* series_dev, int B, int T, int hidden, const float* w_ih,
                            const float* w_hh, const float* b_ih, const float* b_hh, float* h_out_dev,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAU_B200_H_ */
