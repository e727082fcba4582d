// plan.h -- the layer graph of one (variant, B, H, W, T, mode, precision) instance and its executor.
// Mirrors the data flow of reference src/model.py:261-292 (U-Net) and :123-193 (U-Net++) on NHWC
// buffers where every torch.cat is a channel slice of a pre-allocated level buffer.
#pragma once
#include <functional>
#include <string>
#include <vector>
#include "../../include/mau_b200.h"
#include "common.h"
#include "conv_ffma.h"
#include "conv_tc.h"
#include "ops.h"

namespace mau {

struct StateInfo {
  std::string name;
  long long numel = 0;
  int role = 0;   // 0 used param, 1 unused param, 2 running stat, 3 counter
};

struct TRef {      // a channel slice of an activation buffer
  int buf = -1, c0 = 0, C = 0;
};

struct Buf {
  std::string name;
  void* ptr = nullptr;    // activation
  void* gptr = nullptr;   // gradient twin (training, allocated lazily)
  int B = 0, H = 0, W = 0, C = 0, cs = 0;   // B = 1 for batch-invariant buffers of a shared-maps plan
  size_t bytes = 0;
  std::vector<char> ginit;   // per 8 channels: gradient slice already holds a contribution
};

struct Ctx {
  void* const* state = nullptr;
  void* const* grads = nullptr;
  const float* maps = nullptr;
  const void* maps_staged = nullptr;   // alternative input: maps already as NHWC in the plan's activation dtype, channel stride round_up(C, 8)
  const float* series = nullptr;
  const float* md = nullptr;
  float* out = nullptr;
  const float* gout = nullptr;
  cudaStream_t st = nullptr;
  const float* f(int i) const { return static_cast<const float*>(state[i]); }
  float* fm(int i) const { return static_cast<float*>(state[i]); }
  float* g(int i) const { return grads ? static_cast<float*>(grads[i]) : nullptr; }
};

struct Op {
  std::string name;
  std::function<int(Ctx&)> run;
  int grad_first = -1, grad_last = -1;   // state gradients that are final after this (backward) op
};

struct ConvLayer {
  std::string name;
  int Cin = 0, Cout = 0, H = 0, W = 0;
  int iw = -1, ib = -1, igamma = -1, ibeta = -1, irm = -1, irv = -1, inbt = -1;
  int in_buf = -1, nseg = 0, seg_start[4] = {0, 0, 0, 0}, seg_len[4] = {0, 0, 0, 0};
  int Kp = 0;
  TRef out;                 // y = relu(bn(conv(x)))
  int zbuf = -1;            // training: pre-BN conv output (later overwritten by dz)
  bool input_needs_grad = true;
  double flops = 0;         // dense reference FLOPs (cfg.batch tiles)
  int B = 0;                // tiles this layer actually runs on (1 for the shared encoder of a sweep plan)
  int head_w = -1, head_b = -1, head_oc = 0, head_tanh = 0;   // eval: 1x1 head fused into this conv's epilogue
  // device scratch
  int* kmap = nullptr;
  void* wpack = nullptr;
  float* scale = nullptr; float* shift = nullptr; float* mean = nullptr; float* rstd = nullptr;
  double* sums = nullptr;   // [2C] stats / backward sums
  double* dbsum = nullptr;  // [C]
  double* sums_local = nullptr;   // [2C] this rank's backward sums (dgamma / dbeta) when the BN sums are all-reduced
  void* wpack_d[4] = {nullptr, nullptr, nullptr, nullptr};
  int Kd = 0;
  ConvTcOp tc; ConvFfmaParams ff;
  ConvTcOp tc_d[4]; ConvFfmaParams ff_d[4];
  WgradTcOp wg[4];
  int wg_swap = 0;          // wgrad v2 orientation (see wgrad_tc.cu)
  int emb_seg = -1;         // input segment that is constant over space (U-Net++ embedding planes): backward in closed form
  cudaEvent_t ev_pack = nullptr;   // training: this layer's weight pack (launched ahead on the second stream) is complete
};

class Plan {
 public:
  mau_config cfg;
  bool dry = false;          // describe-only: no CUDA calls
  int dt = DT_BF16;
  bool use_tc = true;
  int conv_mode = MODE_ROW3;
  std::vector<StateInfo> state;
  std::vector<Buf> bufs;
  std::vector<ConvLayer*> layers;
  std::vector<Op> fwd, bwd;
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  double fwd_flops = 0, bwd_flops = 0;
  bool forward_done = false;
  mau_grad_ready_fn hook = nullptr; void* hook_user = nullptr;
  mau_stats_sync_fn sync_fn = nullptr; void* sync_user = nullptr; int sync_world = 1;   // SyncBN across ranks
  bool profiling = false;
  std::vector<std::pair<std::string, float>> prof;
  // kernel-level timers (profiling only): CUDA events right around the conv / dgrad / wgrad launches, reported as
  // "k:<layer>:<fwd|dgrad|wgrad>" entries next to the per-op entries
  struct KTimer { std::string name; cudaEvent_t e0, e1; };
  std::vector<KTimer> ktimers;
  void kbegin(const Ctx& c, const std::string& name);
  void kend(const Ctx& c);
  std::vector<int> counter_idx; long long** counters_dev = nullptr;   // num_batches_tracked state indices / device pointer table
  std::string describe_json;
  std::vector<std::string> op_desc;   // JSON objects of the non-conv ops (pool / up / bcast / emb / head), for describe
  void note_op(const char* kind, const TRef* src, const TRef* dst, const std::string& extra = "");
  std::vector<void*> last_state; const float* last_series = nullptr; const float* last_md = nullptr;
  // eval-mode weight cache: packed weights / folded BN are reused while the caller's state is unchanged
  unsigned long long state_version = 0, packed_version = 0;
  std::vector<void*> packed_state;
  bool skip_pack = false;
  bool packs_ahead = false;  // this forward's weight packs were launched ahead on the second stream (training)
  bool shared = false;       // MAU_FLAG_SHARED_MAPS: encoder and LSTM run once, results broadcast over the batch
  double exec_flops = 0;     // FLOPs the convolution kernels execute per forward (== fwd conv FLOPs unless shared)
  double exec_bwd_flops = 0; // FLOPs the dgrad + wgrad kernels execute per backward (dense minus closed-form segments)

  ~Plan();
  int build();
  int run_forward(Ctx& c);
  int run_backward(Ctx& c);

 private:
  // ---- construction helpers
  int add_state(const std::string& name, long long numel, int role);
  int new_buf(const std::string& name, int H, int W, int C, int B = -1);   // B < 0: cfg.batch
  void* alloc(size_t bytes);
  View view(const TRef& t) const;
  View whole(int buf) const;
  int gview(const TRef& t, View* out);        // gradient twin view (allocates lazily)
  int gcontrib(const TRef& t, int* accumulate);   // first contribution stores, later ones accumulate
  struct BlockIdx { int c1w, c1b, g1, b1, rm1, rv1, n1, c2w, c2b, g2, b2, rm2, rv2, n2; };
  BlockIdx add_block_state(const std::string& prefix, int cin, int cmid, int cout);
  int add_encoder_state(int* lstm0, int* fc0, int* mlp0);
  ConvLayer* add_conv(const std::string& name, int iw0, int in_buf, int nseg, const int* seg_start,
                      const int* seg_len, TRef out, bool input_needs_grad);
  int emit_conv_fwd(ConvLayer* L);
  int emit_conv_bwd(ConvLayer* L);
  int add_vgg(const std::string& name, const BlockIdx& bi, int in_buf, int nseg, const int* seg_start,
              const int* seg_len, int cmid, TRef out, bool input_needs_grad, std::vector<ConvLayer*>* made);
  int make_bilinear(int Hin, int Win, int Hout, int Wout, BilinearTables* t);
  int build_unet();
  int build_unetpp();
  int build_encoders(int lstm0, int fc0, int mlp0);                                   // launch on the side stream
  int build_embed_broadcast(const TRef* t_dst, int n_t, const TRef* m_dst, int n_m);  // join + broadcast into the slices
  int side_fork(Ctx& c);
  int side_join(Ctx& c);
 public:
  // Backward overlap: the weight-gradient launches of a layer (workspace memset, wgrad, transpose, closed-form embedding
  // columns) run on a plan-owned second stream, forked after that layer's BatchNorm backward; the caller's stream goes
  // on with the data gradient and the NEXT layer's bandwidth-bound BatchNorm / pool / bilinear backward, which then
  // share the SMs with the tensor-bound wgrad kernel instead of waiting for it.  Joined at the end of backward (and
  // by mau_plan_wait_backward_streams for a data-parallel bucket).  Off while profiling (per-op times stay additive).
  bool conv_stats = true;           // training forward: BatchNorm statistics accumulated inside the convolution kernel
  bool overlap_wgrad = true;
  cudaStream_t wst = nullptr; cudaEvent_t ev_w_fork = nullptr, ev_w_join = nullptr; bool w_pending = false;
  int w_fork(Ctx& c, cudaStream_t* out);
  int w_join(cudaStream_t waiter);
 private:
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool side_pending = false;
  int enc_lstm0 = -1, enc_fc0 = -1, enc_mlp0 = -1;
  int next_emb_seg = -1;            // consumed by the next add_conv
  bool emb_direct = false;          // embedding gradient accumulated straight into demb by the decoder nodes (embgrad.cu)
  float* embgrad_scratch = nullptr;
  std::vector<std::function<int()>> bwd_makers;   // run in reverse to emit backward ops
  // encoder scratch
  float* emb = nullptr; float* demb = nullptr; float* hidden = nullptr; float* hlast = nullptr;
  float* dhlast = nullptr; float* lstm_save = nullptr;
  int emb_dim = 0;
  double* sums_all = nullptr; size_t sums_bytes = 0;   // training: every layer's BatchNorm sums, one block, zeroed once per pass
  float* wgrad_ws = nullptr;   // fp32 [9][Cout][Cin] scratch of the tcgen05 weight-gradient kernel
};

}  // namespace mau
