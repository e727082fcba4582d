// plan.cu -- builds and runs the layer graph (see plan.h).
#include "plan.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <sstream>

namespace mau {

static constexpr float kBnEps = 1e-5f;       // nn.BatchNorm2d default (reference src/model.py:13)
static constexpr float kBnMomentum = 0.1f;

void Plan::kbegin(const Ctx& c, const std::string& name) {
  if (!profiling) return;
  KTimer t; t.name = name;
  cudaEventCreate(&t.e0); cudaEventCreate(&t.e1);
  cudaEventRecord(t.e0, c.st);
  ktimers.push_back(t);
}
void Plan::kend(const Ctx& c) {
  if (!profiling || ktimers.empty()) return;
  cudaEventRecord(ktimers.back().e1, c.st);
}

Plan::~Plan() {
  if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); }
  if (wst) { cudaStreamSynchronize(wst); cudaStreamDestroy(wst); cudaEventDestroy(ev_w_fork); cudaEventDestroy(ev_w_join); }
  if (!dry) {
    for (void* p : allocs) cudaFree(p);
  }
  for (ConvLayer* L : layers) { if (L->ev_pack) cudaEventDestroy(L->ev_pack); delete L; }
}

void* Plan::alloc(size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  ws_bytes += bytes;
  if (dry) return nullptr;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { fail("cudaMalloc of %zu bytes failed", bytes); return nullptr; }
  cudaMemset(p, 0, bytes);
  allocs.push_back(p);
  return p;
}

void Plan::note_op(const char* kind, const TRef* src, const TRef* dst, const std::string& extra) {
  std::ostringstream o;
  o << "{\"kind\":\"" << kind << "\"";
  if (src) o << ",\"src\":[\"" << bufs[src->buf].name << "\"," << src->c0 << "," << src->C << "]";
  if (dst) o << ",\"dst\":[\"" << bufs[dst->buf].name << "\"," << dst->c0 << "," << dst->C << "]";
  if (!extra.empty()) o << "," << extra;
  o << "}";
  op_desc.push_back(o.str());
}

int Plan::add_state(const std::string& name, long long numel, int role) {
  state.push_back(StateInfo{name, numel, role});
  return (int)state.size() - 1;
}

int Plan::new_buf(const std::string& name, int H, int W, int C, int B) {
  Buf b;
  b.name = name; b.B = B < 0 ? cfg.batch : B; b.H = H; b.W = W; b.C = C; b.cs = round_up(C, 8);
  b.bytes = (size_t)b.B * H * W * b.cs * dtype_size(dt);
  b.ptr = alloc(b.bytes);
  b.ginit.assign(b.cs / 8, 0);
  bufs.push_back(b);
  return (int)bufs.size() - 1;
}

View Plan::view(const TRef& t) const {
  const Buf& b = bufs[t.buf];
  View v; v.ptr = b.ptr; v.B = b.B; v.H = b.H; v.W = b.W; v.cs = b.cs; v.c0 = t.c0; v.C = t.C;
  return v;
}
View Plan::whole(int buf) const { return view(TRef{buf, 0, bufs[buf].C}); }

int Plan::gview(const TRef& t, View* out) {
  Buf& b = bufs[t.buf];
  if (!b.gptr && !dry) {
    b.gptr = alloc(b.bytes);
    if (!b.gptr) return -1;
  } else if (dry && !b.gptr) {
    ws_bytes += b.bytes;
    b.gptr = reinterpret_cast<void*>(1);
  }
  *out = view(t);
  out->ptr = dry ? nullptr : b.gptr;
  return 0;
}

int Plan::gcontrib(const TRef& t, int* accumulate) {
  Buf& b = bufs[t.buf];
  int n_init = 0, n = 0;
  for (int g = t.c0 / 8; g < (t.c0 + t.C + 7) / 8; ++g, ++n) n_init += b.ginit[g] ? 1 : 0;
  if (n_init != 0 && n_init != n)
    return fail("gradient slice [%d,%d) of %s is partially initialised", t.c0, t.c0 + t.C, b.name.c_str());
  *accumulate = n_init == n ? 1 : 0;
  for (int g = t.c0 / 8; g < (t.c0 + t.C + 7) / 8; ++g) b.ginit[g] = 1;
  return 0;
}

Plan::BlockIdx Plan::add_block_state(const std::string& p, int cin, int cmid, int cout) {
  BlockIdx b;
  b.c1w = add_state(p + ".conv1.weight", (long long)cmid * cin * 9, 0);
  b.c1b = add_state(p + ".conv1.bias", cmid, 0);
  b.g1 = add_state(p + ".bn1.weight", cmid, 0);
  b.b1 = add_state(p + ".bn1.bias", cmid, 0);
  b.rm1 = add_state(p + ".bn1.running_mean", cmid, 2);
  b.rv1 = add_state(p + ".bn1.running_var", cmid, 2);
  b.n1 = add_state(p + ".bn1.num_batches_tracked", 1, 3);
  b.c2w = add_state(p + ".conv2.weight", (long long)cout * cmid * 9, 0);
  b.c2b = add_state(p + ".conv2.bias", cout, 0);
  b.g2 = add_state(p + ".bn2.weight", cout, 0);
  b.b2 = add_state(p + ".bn2.bias", cout, 0);
  b.rm2 = add_state(p + ".bn2.running_mean", cout, 2);
  b.rv2 = add_state(p + ".bn2.running_var", cout, 2);
  b.n2 = add_state(p + ".bn2.num_batches_tracked", 1, 3);
  return b;
}

int Plan::add_encoder_state(int* lstm0, int* fc0, int* mlp0) {
  const int Hd = cfg.lstm_dim;
  const int tr = cfg.temporal_embeddings ? 0 : 1, mr = cfg.metadata_embeddings ? 0 : 1;
  *lstm0 = add_state("model.temporal_encoder.lstm.weight_ih_l0", 4LL * Hd, tr);
  add_state("model.temporal_encoder.lstm.weight_hh_l0", 4LL * Hd * Hd, tr);
  add_state("model.temporal_encoder.lstm.bias_ih_l0", 4LL * Hd, tr);
  add_state("model.temporal_encoder.lstm.bias_hh_l0", 4LL * Hd, tr);
  *fc0 = add_state("model.temporal_encoder.fc.weight", (long long)cfg.temporal_dim * Hd, tr);
  add_state("model.temporal_encoder.fc.bias", cfg.temporal_dim, tr);
  *mlp0 = add_state("model.meta_encoder.fc.0.weight", 32LL * cfg.meta_features, mr);
  add_state("model.meta_encoder.fc.0.bias", 32, mr);
  add_state("model.meta_encoder.fc.2.weight", (long long)cfg.meta_dim * 32, mr);
  add_state("model.meta_encoder.fc.2.bias", cfg.meta_dim, mr);
  return 0;
}

// ------------------------------------------------------------------------------------------ conv + BN + ReLU
ConvLayer* Plan::add_conv(const std::string& name, int iw0, int in_buf, int nseg, const int* seg_start,
                          const int* seg_len, TRef out, bool input_needs_grad) {
  ConvLayer* L = new ConvLayer();
  layers.push_back(L);
  L->name = name;
  L->iw = iw0; L->ib = iw0 + 1; L->igamma = iw0 + 2; L->ibeta = iw0 + 3; L->irm = iw0 + 4; L->irv = iw0 + 5;
  L->inbt = iw0 + 6;
  L->in_buf = in_buf; L->nseg = nseg; L->out = out;
  L->H = bufs[in_buf].H; L->W = bufs[in_buf].W;
  L->Cout = out.C;
  L->input_needs_grad = input_needs_grad;
  L->emb_seg = next_emb_seg; next_emb_seg = -1;
  std::vector<int> kmap;
  for (int s = 0; s < nseg; ++s) {
    L->seg_start[s] = seg_start[s]; L->seg_len[s] = seg_len[s];
    for (int i = 0; i < round_up(seg_len[s], 64); ++i) kmap.push_back(i < seg_len[s] ? L->Cin + i : -1);
    L->Cin += seg_len[s];
  }
  L->Kp = (int)kmap.size();
  L->B = bufs[in_buf].B;
  if (bufs[out.buf].B != L->B) { fail("layer %s: input and output batch differ", name.c_str()); return nullptr; }
  L->flops = 2.0 * 9.0 * L->Cin * L->Cout * (double)L->H * L->W * cfg.batch;
  exec_flops += L->flops / cfg.batch * L->B;
  if ((long long)L->Cout * L->Cin * 9 != state[iw0].numel) {
    fail("layer %s: weight numel mismatch (%d x %d)", name.c_str(), L->Cout, L->Cin);
    return nullptr;
  }
  const int C = L->Cout;
  L->kmap = static_cast<int*>(alloc(sizeof(int) * kmap.size()));
  L->wpack = alloc(use_tc ? (size_t)9 * C * L->Kp * 2 : (size_t)9 * C * L->Kp * 4);
  L->scale = static_cast<float*>(alloc(sizeof(float) * C));
  L->shift = static_cast<float*>(alloc(sizeof(float) * C));
  L->mean = static_cast<float*>(alloc(sizeof(float) * C));
  L->rstd = static_cast<float*>(alloc(sizeof(float) * C));
  if (!cfg.training) L->sums = static_cast<double*>(alloc(sizeof(double) * 2 * C));    // training: carved from sums_all in build()
  L->dbsum = static_cast<double*>(alloc(sizeof(double) * C));
  if (cfg.training) L->sums_local = static_cast<double*>(alloc(sizeof(double) * 2 * C));
  if (cfg.training) L->zbuf = new_buf(name + ".z", L->H, L->W, C, L->B);
  if (dry) return L;
  if (!L->kmap || !L->wpack || !L->dbsum) return nullptr;
  if (cudaMemcpy(L->kmap, kmap.data(), sizeof(int) * kmap.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    fail("kmap upload failed");
    return nullptr;
  }
  const View xin = whole(in_buf);
  const View dst = cfg.training ? whole(L->zbuf) : view(out);
  int rc;
  if (use_tc)
    rc = conv_tc_prepare(&L->tc, xin, nseg, L->seg_start, L->seg_len, L->wpack, L->Kp, C, dst, conv_mode, nullptr,
                         nullptr, 0, 0, (cfg.flags >> 8) & 1);
  else
    rc = conv_ffma_prepare(&L->ff, xin, nseg, L->seg_start, L->seg_len, static_cast<const float*>(L->wpack), L->Kp,
                           C, dst, nullptr, nullptr, 0, 0);
  if (rc) return nullptr;
  return L;
}

int Plan::emit_conv_fwd(ConvLayer* L) {
  fwd_flops += L->flops;
  bwd_flops += L->flops * (L->input_needs_grad ? 2.0 : 1.0);
  if (dry) return 0;
  Op op;
  op.name = L->name;
  const bool training = cfg.training != 0;
  const long long count = (long long)L->B * L->H * L->W;
  op.run = [this, L, training, count](Ctx& c) -> int {
    const int C = L->Cout;
    if (packs_ahead) {
      MAU_CUDA(cudaStreamWaitEvent(c.st, L->ev_pack, 0));      // packed by run_forward on the second stream
    } else if (!skip_pack) {
      if (use_tc) MAU_TRY(conv_tc_pack_fwd(c.f(L->iw), C, L->Cin, L->kmap, L->Kp, L->wpack, c.st));
      else MAU_TRY(conv_ffma_pack_fwd(c.f(L->iw), C, L->Cin, L->kmap, L->Kp, static_cast<float*>(L->wpack), c.st));
    }
    const float* scale = nullptr; const float* shift = c.f(L->ib); int relu = 0;
    if (!training) {
      if (!skip_pack)
        MAU_TRY(op_bn_fold_eval(c.f(L->igamma), c.f(L->ibeta), c.f(L->irm), c.f(L->irv), c.f(L->ib), C, kBnEps,
                                L->scale, L->shift, c.st));
      scale = L->scale; shift = L->shift; relu = 1;
    }
    // training, tcgen05 halo kernels: the BatchNorm statistics are accumulated by the convolution itself (idle warps read
    // each staged output tile), so z is not read again by a separate bn_stats pass
    const bool fused_stats = training && use_tc && conv_mode == MODE_HALO && conv_stats;
    if (use_tc) {
      L->tc.p.stats = fused_stats ? L->sums : nullptr;
      L->tc.p.scale = scale; L->tc.p.shift = shift; L->tc.p.relu = relu;
      if (L->head_w >= 0) {       // fused 1x1 head: the activation itself is never stored
        L->tc.p.head_w = c.f(L->head_w); L->tc.p.head_b = c.f(L->head_b); L->tc.p.head_out = c.out;
        L->tc.p.head_oc = L->head_oc; L->tc.p.head_tanh = L->head_tanh; L->tc.p.store_y = 0;
      }
      kbegin(c, "k:" + L->name + ":fwd");
      MAU_TRY(conv_tc_launch(L->tc, c.st));
      kend(c);
    } else {
      L->ff.scale = scale; L->ff.shift = shift; L->ff.relu = relu;
      kbegin(c, "k:" + L->name + ":fwd");
      MAU_TRY(conv_ffma_launch(dt, L->ff, L->B, c.st));
      kend(c);
    }
    if (training) {
      const View z = whole(L->zbuf);
      // (statistics fused into the conv epilogue were measured and dropped: the column-sum pass over the staged
      //  tile lengthens the epilogue of the narrow layers by more than this separate 63 %-of-HBM-peak pass costs)
      // (a single cooperative launch -- statistics, grid barrier, apply over the same block ranges in reverse so that the
      //  second read of z hits L2 -- was measured and dropped: 8.41 vs 8.16 ms per step, profiles/r02_training_step.md)
      if (!fused_stats) MAU_TRY(op_bn_stats(dt, z, L->sums, c.st));
      if (sync_fn) sync_fn(sync_user, L->sums, 2 * C);       // SyncBN: sum / sum-of-squares over all ranks
      MAU_TRY(op_bn_finalize_apply_relu(dt, z, L->sums, count * sync_world, c.f(L->igamma), c.f(L->ibeta), kBnEps, kBnMomentum,
                                        c.fm(L->irm), c.fm(L->irv), L->scale, L->shift, L->mean, L->rstd, view(L->out), c.st));
    }
    return 0;
  };
  fwd.push_back(op);
  if (training) {
    counter_idx.push_back(L->inbt);     // state index of num_batches_tracked, resolved to a pointer at run time
    bwd_makers.push_back([this, L]() { return emit_conv_bwd(L); });
  }
  return 0;
}

int Plan::emit_conv_bwd(ConvLayer* L) {
  View gy;
  MAU_TRY(gview(L->out, &gy));
  {
    int acc = 0;
    TRef t = L->out;
    // the output gradient must already hold every consumer's contribution
    Buf& b = bufs[t.buf];
    for (int g = t.c0 / 8; g < (t.c0 + t.C + 7) / 8; ++g)
      if (!b.ginit[g]) return fail("backward: gradient of %s was never produced", L->name.c_str());
    (void)acc;
  }
  const int C = L->Cout;
  const View y = view(L->out);
  const View z = whole(L->zbuf);
  const long long count = (long long)L->B * L->H * L->W;
  // weight-gradient and data-gradient launches are prepared now (tensor maps need final pointers)
  int ci_w0 = 0;
  int dacc[4] = {0, 0, 0, 0};
  L->Kd = round_up(C, 64);
  L->wg_swap = (use_tc && wgrad_ws) ? wgrad_tc_pick_swap(C, L->nseg, L->seg_len) : 0;
  if (L->emb_seg >= 0) {
    emb_direct = true;
    if (!embgrad_scratch && !dry) {
      size_t fl = 0;
      for (const ConvLayer* q : layers) fl = std::max(fl, emb_grad_scratch_floats(cfg.batch, q->Cout));
      embgrad_scratch = static_cast<float*>(alloc(sizeof(float) * fl));
      if (!embgrad_scratch) return -1;
    }
  }
  for (int s = 0; s < L->nseg; ++s) {
    if (s == L->emb_seg) { ci_w0 += L->seg_len[s]; continue; }     // no wgrad / dgrad launches for the constant planes
    exec_bwd_flops += 2.0 * 9.0 * L->seg_len[s] * C * (double)L->H * L->W * L->B * (L->input_needs_grad ? 2.0 : 1.0);
    const TRef xin{L->in_buf, L->seg_start[s], L->seg_len[s]};
    if (use_tc) MAU_TRY(wgrad_tc_prepare(&L->wg[s], view(xin), z, ci_w0, L->Cin, wgrad_ws, L->wg_swap));
    if (L->input_needs_grad) {
      View gx;
      MAU_TRY(gview(xin, &gx));
      MAU_TRY(gcontrib(xin, &dacc[s]));
      const int zero = 0;
      if (use_tc && conv_mode == MODE_HALO) {
        // no transposed re-pack: the kernel reads W^T MN-major out of the forward pack (packed column of this segment)
        int col0 = 0;
        for (int q = 0; q < s; ++q) col0 += round_up(L->seg_len[q], 64);
        MAU_TRY(conv_tc_prepare_dgrad(&L->tc_d[s], z, L->wpack, L->Kp, col0, gx, dacc[s]));
      } else if (use_tc) {
        L->wpack_d[s] = alloc((size_t)9 * L->seg_len[s] * L->Kd * 2);
        if (!L->wpack_d[s]) return -1;
        MAU_TRY(conv_tc_prepare(&L->tc_d[s], z, 1, &zero, &C, L->wpack_d[s], L->Kd, L->seg_len[s], gx, conv_mode,
                                nullptr, nullptr, 0, dacc[s], (cfg.flags >> 8) & 1));
      } else {
        L->wpack_d[s] = alloc((size_t)9 * L->seg_len[s] * L->Kd * 4);
        if (!L->wpack_d[s]) return -1;
        MAU_TRY(conv_ffma_prepare(&L->ff_d[s], z, 1, &zero, &C, static_cast<const float*>(L->wpack_d[s]), L->Kd,
                                  L->seg_len[s], gx, nullptr, nullptr, 0, dacc[s]));
      }
    }
    ci_w0 += L->seg_len[s];
  }
  Op op;
  op.name = L->name + ".bwd";
  op.grad_first = L->iw; op.grad_last = L->ibeta;
  op.run = [this, L, gy, y, z, count, C](Ctx& c) -> int {
    MAU_TRY(op_bn_bwd_reduce(dt, gy, z, L->scale, L->shift, L->mean, L->rstd, L->sums, c.st));
    const double* param_sums = L->sums;
    if (sync_fn) {
      // SyncBN: dz needs the sums over every rank's pixels; dgamma / dbeta stay this rank's own sums (the
      // data-parallel gradient average then yields the global-batch gradient, like torch SyncBatchNorm)
      MAU_CUDA(cudaMemcpyAsync(L->sums_local, L->sums, sizeof(double) * 2 * C, cudaMemcpyDeviceToDevice, c.st));
      sync_fn(sync_user, L->sums, 2 * C);
      param_sums = L->sums_local;
    }
    MAU_TRY(op_bn_bwd_apply(dt, gy, z, L->scale, L->shift, c.f(L->igamma), L->mean, L->rstd, L->sums,
                            count * sync_world, z, param_sums, c.g(L->igamma), c.g(L->ibeta), c.g(L->ib), c.st));
    float* dw = c.g(L->iw);
    int ci_w0 = 0;
    const bool ws_path = use_tc && wgrad_ws != nullptr;      // v2 wgrad: reduce into the workspace, then transpose
    // The second stream forks right after this layer's BatchNorm backward: data gradient (on the dependency chain, enqueued
    // first) and weight gradient are runnable at once and share the SMs.  Forking BEHIND the data gradient instead
    // (MAU_WGRAD_FORK_LATE=1: the persistent wgrad CTAs then never hold SMs the chain's tensor kernel waits for) was measured
    // slower, 7.14-7.18 vs 6.98 ms per step on the same box: the weight gradients pile up and leave a tail after the chain.
    static const bool fork_early = [] { const char* e = getenv("MAU_WGRAD_FORK_LATE"); return !(e && atoi(e) != 0); }();
    cudaStream_t ws = c.st;                                  // weight-gradient stream (== c.st when the overlap is off)
    if (fork_early) MAU_TRY(w_fork(c, &ws));
    if (L->input_needs_grad) {
      int ci0 = 0;
      for (int s = 0; s < L->nseg; ++s) {
        if (s == L->emb_seg) { ci0 += L->seg_len[s]; continue; }
        if (use_tc) {
          if (L->wpack_d[s])
            MAU_TRY(conv_tc_pack_dgrad(c.f(L->iw), C, L->Cin, ci0, L->seg_len[s], L->Kd, L->wpack_d[s], c.st));
          kbegin(c, "k:" + L->name + ":dgrad");
          MAU_TRY(conv_tc_launch(L->tc_d[s], c.st));
          kend(c);
        } else {
          MAU_TRY(conv_ffma_pack_dgrad(c.f(L->iw), C, L->Cin, ci0, L->seg_len[s], L->Kd,
                                       static_cast<float*>(L->wpack_d[s]), c.st));
          MAU_TRY(conv_ffma_launch(dt, L->ff_d[s], L->B, c.st));
        }
        ci0 += L->seg_len[s];
      }
    }
    if (!fork_early) MAU_TRY(w_fork(c, &ws));
    Ctx cw = c; cw.st = ws;
    if (dw && ws_path) MAU_CUDA(cudaMemsetAsync(wgrad_ws, 0, sizeof(float) * wgrad_tc_workspace_floats(C, L->Cin, L->wg_swap), ws));
    else if (dw) MAU_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * L->Cin * 9, ws));
    int emb_ci0 = -1;
    for (int s = 0; s < L->nseg; ++s) {
      if (s == L->emb_seg) { emb_ci0 = ci_w0; ci_w0 += L->seg_len[s]; continue; }
      const TRef xin{L->in_buf, L->seg_start[s], L->seg_len[s]};
      if (dw) {
        kbegin(cw, "k:" + L->name + ":wgrad");
        if (use_tc) MAU_TRY(wgrad_tc_launch(L->wg[s], dw, ws));
        else MAU_TRY(wgrad_ffma_launch(dt, view(xin), z, ci_w0, L->Cin, dw, 1, ws));
        kend(cw);
      }
      ci_w0 += L->seg_len[s];
    }
    if (dw && ws_path) MAU_TRY(wgrad_tc_finalize(wgrad_ws, L->wg_swap, C, L->Cin, dw, ws));
    if (emb_ci0 >= 0)      // dW of the constant planes' columns and their contribution to d emb, from nine sums of dz per image
      MAU_TRY(op_emb_segment_grad(dt, z, c.f(L->iw), L->Cin, emb_ci0, L->seg_len[L->emb_seg], emb, emb_dim, dw, demb,
                                  embgrad_scratch, ws));
    return 0;
  };
  bwd.push_back(op);
  return 0;
}

int Plan::add_vgg(const std::string& name, const BlockIdx& bi, int in_buf, int nseg, const int* seg_start,
                  const int* seg_len, int cmid, TRef out, bool input_needs_grad, std::vector<ConvLayer*>* made) {
  const int mid = new_buf(name + ".mid", bufs[in_buf].H, bufs[in_buf].W, cmid, bufs[in_buf].B);
  ConvLayer* a = add_conv(name + ".conv1", bi.c1w, in_buf, nseg, seg_start, seg_len, TRef{mid, 0, cmid},
                          input_needs_grad);
  if (!a) return -1;
  MAU_TRY(emit_conv_fwd(a));
  const int zero = 0;
  ConvLayer* b = add_conv(name + ".conv2", bi.c2w, mid, 1, &zero, &cmid, out, true);
  if (!b) return -1;
  MAU_TRY(emit_conv_fwd(b));
  if (made) { made->push_back(a); made->push_back(b); }
  return 0;
}

int Plan::make_bilinear(int Hin, int Win, int Hout, int Wout, BilinearTables* t) {
  t->Hin = Hin; t->Win = Win; t->Hout = Hout; t->Wout = Wout;
  if (dry) return 0;
  BilinearHost hy, hx;
  bilinear_axis_tables(Hin, Hout, &hy);
  bilinear_axis_tables(Win, Wout, &hx);
  auto up_i = [&](const std::vector<int>& v) -> int* {
    int* d = static_cast<int*>(alloc(sizeof(int) * std::max<size_t>(1, v.size())));
    if (d) cudaMemcpy(d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice);
    return d;
  };
  auto up_f = [&](const std::vector<float>& v) -> float* {
    float* d = static_cast<float*>(alloc(sizeof(float) * std::max<size_t>(1, v.size())));
    if (d) cudaMemcpy(d, v.data(), sizeof(float) * v.size(), cudaMemcpyHostToDevice);
    return d;
  };
  t->y0 = up_i(hy.i0); t->y1 = up_i(hy.i1); t->ly = up_f(hy.l);
  t->x0 = up_i(hx.i0); t->x1 = up_i(hx.i1); t->lx = up_f(hx.l);
  t->ty_off = up_i(hy.t_off); t->ty_idx = up_i(hy.t_idx); t->ty_w = up_f(hy.t_w);
  t->tx_off = up_i(hx.t_off); t->tx_idx = up_i(hx.t_idx); t->tx_w = up_f(hx.t_w);
  t->max_fan_w = hx.max_fan;
  if (!t->tx_w || !t->ty_w) return -1;
  return 0;
}

// ------------------------------------------------------------------------------------------ encoders
// The LSTM is 828 strictly sequential steps (~0.6 ms forward, ~1.5 ms BPTT) on a handful of SMs.  It runs on a
// plan-owned side stream: forward, it is launched first and joined right before the first consumer of the
// embedding (the bottleneck of the U-Net, node x0_1 of the U-Net++), so it overlaps the encoder convolutions;
// backward, the embedding gradient is complete right after that consumer's backward, the BPTT then overlaps the
// encoder convolutions' backward and is joined at the end of the pass (where its gradients are reported ready).
int Plan::side_fork(Ctx& c) {
  if (!side) {
    MAU_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    MAU_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    MAU_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  MAU_CUDA(cudaEventRecord(ev_fork, c.st));
  MAU_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  return 0;
}
int Plan::side_join(Ctx& c) {
  if (!side || !side_pending) return 0;
  MAU_CUDA(cudaStreamWaitEvent(c.st, ev_join, 0));
  side_pending = false;
  return 0;
}

int Plan::w_fork(Ctx& c, cudaStream_t* out) {
  *out = c.st;
  if (!overlap_wgrad || profiling) return 0;
  if (!wst) {
    MAU_CUDA(cudaStreamCreateWithFlags(&wst, cudaStreamNonBlocking));
    MAU_CUDA(cudaEventCreateWithFlags(&ev_w_fork, cudaEventDisableTiming));
    MAU_CUDA(cudaEventCreateWithFlags(&ev_w_join, cudaEventDisableTiming));
  }
  MAU_CUDA(cudaEventRecord(ev_w_fork, c.st));
  MAU_CUDA(cudaStreamWaitEvent(wst, ev_w_fork, 0));
  w_pending = true;
  *out = wst;
  return 0;
}
// `waiter` waits for everything enqueued on the weight-gradient stream so far
int Plan::w_join(cudaStream_t waiter) {
  if (!wst || !w_pending) return 0;
  MAU_CUDA(cudaEventRecord(ev_w_join, wst));
  MAU_CUDA(cudaStreamWaitEvent(waiter, ev_w_join, 0));
  return 0;
}

int Plan::build_encoders(int lstm0, int fc0, int mlp0) {
  const bool te = cfg.temporal_embeddings != 0, me = cfg.metadata_embeddings != 0;
  if (!te && !me) return 0;
  const int B = cfg.batch, Hd = cfg.lstm_dim, td = cfg.temporal_dim, md = cfg.meta_dim, T = cfg.seq_len;
  const int t_off = 0, m_off = te ? td : 0;
  emb_dim = (te ? td : 0) + (me ? md : 0);
  emb = static_cast<float*>(alloc(sizeof(float) * B * emb_dim));
  demb = static_cast<float*>(alloc(sizeof(float) * B * emb_dim));
  hidden = static_cast<float*>(alloc(sizeof(float) * B * 32));
  hlast = static_cast<float*>(alloc(sizeof(float) * B * Hd));
  dhlast = static_cast<float*>(alloc(sizeof(float) * B * Hd));
  if (te && cfg.training) lstm_save = static_cast<float*>(alloc(sizeof(float) * lstm_save_floats(B, T, Hd)));
  if (te) fwd_flops += B * (2.0 * 4 * Hd * (Hd + 1) * T + 2.0 * Hd * td);
  if (me) fwd_flops += B * (2.0 * cfg.meta_features * 32 + 2.0 * 32 * md);
  enc_lstm0 = lstm0; enc_fc0 = fc0; enc_mlp0 = mlp0;
  if (dry) return 0;
  if (te && T < 1) return fail("temporal embeddings enabled but temp_series is empty");
  const int Bt = shared ? 1 : B;                 // shared sweep: one LSTM run, row 0 broadcast to every tile
  Op op;
  op.name = "encoders";                          // launch only: the work runs on the side stream
  op.run = [=](Ctx& c) -> int {
    if (te && !c.series) return fail("temp_series is required when temporal embeddings are enabled");
    if (me && !c.md) return fail("metadata is required when metadata embeddings are enabled");
    MAU_TRY(side_fork(c));
    if (te) {
      MAU_TRY(op_lstm_fwd(c.series, Bt, T, Hd, c.f(lstm0), c.f(lstm0 + 1), c.f(lstm0 + 2), c.f(lstm0 + 3), hlast,
                          lstm_save, side));
      MAU_TRY(op_linear_fwd(hlast, Bt, Hd, c.f(fc0), c.f(fc0 + 1), td, emb + t_off, emb_dim, side));
    }
    if (me)
      MAU_TRY(op_mlp_fwd(c.md, B, cfg.meta_features, c.f(mlp0), c.f(mlp0 + 1), c.f(mlp0 + 2), c.f(mlp0 + 3), md,
                         hidden, emb + m_off, emb_dim, side));
    MAU_CUDA(cudaEventRecord(ev_join, side));
    side_pending = true;
    return 0;
  };
  fwd.push_back(op);
  if (cfg.training)
    bwd_makers.push_back([=]() -> int {          // runs LAST in backward: join the side stream, gradients are final
      Op b;
      b.name = "encoders.bwd.join";
      b.grad_first = std::min(lstm0, mlp0); b.grad_last = std::max(fc0 + 1, mlp0 + 3);
      b.run = [=](Ctx& c) -> int { return side_join(c); };
      bwd.push_back(b);
      return 0;
    });
  return 0;
}

// placed right before the first consumer of the embedding slices
int Plan::build_embed_broadcast(const TRef* t_dst, int n_t, const TRef* m_dst, int n_m) {
  const bool te = cfg.temporal_embeddings != 0, me = cfg.metadata_embeddings != 0;
  if (te) for (int i = 0; i < n_t; ++i) note_op("emb", nullptr, &t_dst[i], "\"which\":\"temporal\"");
  if (me) for (int i = 0; i < n_m; ++i) note_op("emb", nullptr, &m_dst[i], "\"which\":\"metadata\"");
  if ((!te && !me) || dry) return 0;
  const int B = cfg.batch, Hd = cfg.lstm_dim, td = cfg.temporal_dim, md = cfg.meta_dim, T = cfg.seq_len;
  const int t_off = 0, m_off = te ? td : 0;
  const int t_stride = shared ? 0 : emb_dim;
  const int lstm0 = enc_lstm0, fc0 = enc_fc0, mlp0 = enc_mlp0;
  std::vector<View> tv, mv;
  for (int i = 0; i < n_t; ++i) tv.push_back(view(t_dst[i]));
  for (int i = 0; i < n_m; ++i) mv.push_back(view(m_dst[i]));
  Op op;
  op.name = "emb.bcast";
  op.run = [=](Ctx& c) -> int {
    MAU_TRY(side_join(c));
    if (te) for (const View& v : tv) MAU_TRY(op_embed_broadcast(dt, emb + t_off, t_stride, v, c.st));
    if (me) for (const View& v : mv) MAU_TRY(op_embed_broadcast(dt, emb + m_off, emb_dim, v, c.st));
    return 0;
  };
  fwd.push_back(op);
  if (cfg.training) {
    std::vector<TRef> tr(t_dst, t_dst + n_t), mr(m_dst, m_dst + n_m);
    bwd_makers.push_back([=]() -> int {
      std::vector<View> gt, gm;
      for (const TRef& t : tr) { View g; MAU_TRY(gview(t, &g)); gt.push_back(g); }
      for (const TRef& t : mr) { View g; MAU_TRY(gview(t, &g)); gm.push_back(g); }
      Op b;
      b.name = "emb.bwd";                        // reduce on the main stream, encoder backward on the side stream
      b.run = [=](Ctx& c) -> int {
        if (emb_direct) MAU_TRY(w_join(c.st));      // d emb was accumulated by the weight-gradient stream
        if (!emb_direct) {     // (U-Net++: the decoder nodes have already accumulated d emb in closed form)
          MAU_CUDA(cudaMemsetAsync(demb, 0, sizeof(float) * B * emb_dim, c.st));
          if (te) for (const View& g : gt) MAU_TRY(op_embed_reduce(dt, g, demb + t_off, emb_dim, 1, c.st));
          if (me) for (const View& g : gm) MAU_TRY(op_embed_reduce(dt, g, demb + m_off, emb_dim, 1, c.st));
        }
        MAU_TRY(side_fork(c));
        if (te) {
          if (c.g(fc0))
            MAU_TRY(op_linear_bwd(hlast, B, Hd, c.f(fc0), td, demb + t_off, emb_dim, dhlast, c.g(fc0),
                                  c.g(fc0 + 1), side));
          if (c.g(lstm0)) {
            MAU_CUDA(cudaMemsetAsync(c.g(lstm0), 0, sizeof(float) * 4 * Hd, side));
            MAU_CUDA(cudaMemsetAsync(c.g(lstm0 + 1), 0, sizeof(float) * 4 * Hd * Hd, side));
            MAU_CUDA(cudaMemsetAsync(c.g(lstm0 + 2), 0, sizeof(float) * 4 * Hd, side));
            MAU_CUDA(cudaMemsetAsync(c.g(lstm0 + 3), 0, sizeof(float) * 4 * Hd, side));
            MAU_TRY(op_lstm_bwd(c.series, B, T, Hd, c.f(lstm0 + 1), lstm_save, dhlast, c.g(lstm0), c.g(lstm0 + 1),
                                c.g(lstm0 + 2), c.g(lstm0 + 3), nullptr, side));
          }
        }
        if (me && c.g(mlp0))
          MAU_TRY(op_mlp_bwd(c.md, B, cfg.meta_features, c.f(mlp0), c.f(mlp0 + 2), md, hidden, demb + m_off,
                             emb_dim, c.g(mlp0), c.g(mlp0 + 1), c.g(mlp0 + 2), c.g(mlp0 + 3), side));
        MAU_CUDA(cudaEventRecord(ev_join, side));
        side_pending = true;
        return 0;
      };
      bwd.push_back(b);
      return 0;
    });
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ U-Net
int Plan::build_unet() {
  const int B = cfg.batch;
  const int* F = cfg.filters;
  const bool te = cfg.temporal_embeddings != 0, me = cfg.metadata_embeddings != 0;
  const int E = (te ? cfg.temporal_dim : 0) + (me ? cfg.meta_dim : 0);
  int Hs[5], Ws[5];
  Hs[0] = cfg.height; Ws[0] = cfg.width;
  for (int l = 1; l < 5; ++l) { Hs[l] = Hs[l - 1] / 2; Ws[l] = Ws[l - 1] / 2; }
  if (Hs[4] < 1 || Ws[4] < 1) return fail("tile too small for four 2x2 poolings");

  // state_dict order: reference src/model.py:214-241
  int lstm0, fc0, mlp0;
  add_encoder_state(&lstm0, &fc0, &mlp0);
  BlockIdx enc[5], dec[4];
  for (int l = 0; l < 4; ++l)
    enc[l] = add_block_state("model.conv" + std::to_string(l) + "_0", l == 0 ? cfg.spatial_channels : F[l - 1], F[l], F[l]);
  enc[4] = add_block_state("model.conv4_0", F[3] + E, F[4], F[4]);
  for (int l = 3; l >= 0; --l) dec[l] = add_block_state("model.conv" + std::to_string(l) + "_1", F[l] + F[l + 1], F[l], F[l]);
  const int fw = add_state("model.final.weight", (long long)cfg.out_channels * F[0], 0);
  const int fb = add_state("model.final.bias", cfg.out_channels, 0);

  // buffers.  Shared-maps sweep (eval): everything up to the bottleneck input is batch-invariant and lives in
  // one-tile buffers; the skip tensors are replicated over the batch rows of the concat buffers.
  const int Be = shared ? 1 : B;
  const int in0 = new_buf("maps_nhwc", Hs[0], Ws[0], cfg.spatial_channels, Be);
  int cat[4], pooled[5], xdec[4], enc1[4] = {-1, -1, -1, -1};
  for (int l = 0; l < 4; ++l) cat[l] = new_buf("cat" + std::to_string(l), Hs[l], Ws[l], F[l] + F[l + 1]);
  for (int l = 1; l < 4; ++l) pooled[l] = new_buf("pool" + std::to_string(l), Hs[l], Ws[l], F[l - 1], Be);
  if (shared) {
    for (int l = 0; l < 4; ++l) enc1[l] = new_buf("x" + std::to_string(l) + "_0.shared", Hs[l], Ws[l], F[l], 1);
    pooled[4] = new_buf("pool4.shared", Hs[4], Ws[4], F[3], 1);
  }
  const int bott = new_buf("bottleneck_in", Hs[4], Ws[4], F[3] + E);
  const int x4 = new_buf("x4_0", Hs[4], Ws[4], F[4]);
  for (int l = 0; l < 4; ++l) xdec[l] = new_buf("x" + std::to_string(l) + "_1", Hs[l], Ws[l], F[l]);

  // encoders -> bottleneck slices (temporal first, then metadata: src/model.py:248-259)
  {
    MAU_TRY(build_encoders(lstm0, fc0, mlp0));
  }
  // maps -> NHWC
  { const TRef t{in0, 0, cfg.spatial_channels}; note_op("input", nullptr, &t); }
  if (!dry) {
    Op op; op.name = "nchw_to_nhwc";
    const View dst = whole(in0);
    op.run = [=](Ctx& c) -> int {
      if (c.maps_staged) {     // pre-staged NHWC tiles (mau_plan_forward_staged): a copy KERNEL instead of the layout kernel
        // (a cudaMemcpyAsync would queue on the copy engines behind the caller's own H2D prefetches: measured, e2e 7.4k vs 8.0k tiles/s)
        View src = dst;
        src.ptr = const_cast<void*>(c.maps_staged);
        View d8 = dst;
        src.C = d8.C = dst.cs;      // whole 8-channel groups, pad channel included
        return op_copy_slice(dt, src, d8, 0, c.st);
      }
      return op_nchw_to_nhwc(dt, c.maps, Be, cfg.spatial_channels, cfg.height, cfg.width, dst, c.st);
    };
    fwd.push_back(op);
  }
  auto add_bcast = [&](TRef src, TRef dst) -> int {       // one-tile tensor -> every batch row of a slice
    note_op("bcast", &src, &dst);
    if (dry) return 0;
    Op op; op.name = "bcast." + bufs[src.buf].name;
    const View xs = view(src), yd = view(dst);
    op.run = [=](Ctx& c) -> int { return op_broadcast_batch(dt, xs, yd, c.st); };
    fwd.push_back(op);
    return 0;
  };
  auto add_pool = [&](TRef src, TRef dst) -> int {
    note_op("pool", &src, &dst);
    if (dry) return 0;
    Op op; op.name = "pool." + bufs[src.buf].name;
    const View xs = view(src), yd = view(dst);
    op.run = [=](Ctx& c) -> int { return op_maxpool(dt, xs, yd, c.st); };
    fwd.push_back(op);
    if (cfg.training)
      bwd_makers.push_back([=]() -> int {
        View gy, gx;
        MAU_TRY(gview(dst, &gy));
        MAU_TRY(gview(src, &gx));
        int acc = 0;
        MAU_TRY(gcontrib(src, &acc));
        Op b; b.name = "pool.bwd." + bufs[src.buf].name;
        b.run = [=](Ctx& c) -> int { return op_maxpool_bwd(dt, xs, gy, acc ? &gx : nullptr, gx, c.st); };
        bwd.push_back(b);
        return 0;
      });
    return 0;
  };
  // self.up (x2) then _upsample_match (src/model.py:219, 243-246): two stages when 2*h != target
  auto add_up = [&](TRef src, TRef dst) -> int {
    const int hs = bufs[src.buf].H, wsrc = bufs[src.buf].W, ht = bufs[dst.buf].H, wt = bufs[dst.buf].W;
    const bool two = (2 * hs != ht) || (2 * wsrc != wt);
    TRef mid = src;
    if (two) mid = TRef{new_buf("up2x." + bufs[src.buf].name, 2 * hs, 2 * wsrc, src.C), 0, src.C};
    if (two) { note_op("up", &src, &mid); note_op("up", &mid, &dst); }
    else note_op("up", &src, &dst);
    BilinearTables t1, t2;
    if (two) { MAU_TRY(make_bilinear(hs, wsrc, 2 * hs, 2 * wsrc, &t1)); MAU_TRY(make_bilinear(2 * hs, 2 * wsrc, ht, wt, &t2)); }
    else MAU_TRY(make_bilinear(hs, wsrc, ht, wt, &t2));
    if (dry) return 0;
    Op op; op.name = "up." + bufs[src.buf].name;
    const View vs = view(src), vm = view(mid), vd = view(dst);
    op.run = [=](Ctx& c) -> int {
      if (two) MAU_TRY(op_bilinear(dt, vs, vm, t1, c.st));
      return op_bilinear(dt, vm, vd, t2, c.st);
    };
    fwd.push_back(op);
    if (cfg.training)
      bwd_makers.push_back([=]() -> int {
        View gd, gm, gs;
        MAU_TRY(gview(dst, &gd));
        MAU_TRY(gview(src, &gs));
        int acc_s = 0, acc_m = 0;
        if (two) { MAU_TRY(gview(mid, &gm)); MAU_TRY(gcontrib(mid, &acc_m)); }
        MAU_TRY(gcontrib(src, &acc_s));
        Op b; b.name = "up.bwd." + bufs[src.buf].name;
        b.run = [=](Ctx& c) -> int {
          if (two) {
            MAU_TRY(op_bilinear_bwd(dt, gd, gm, t2, 0, c.st));
            return op_bilinear_bwd(dt, gm, gs, t1, acc_s, c.st);
          }
          return op_bilinear_bwd(dt, gd, gs, t2, acc_s, c.st);
        };
        bwd.push_back(b);
        return 0;
      });
    return 0;
  };

  const int zero = 0;
  std::vector<ConvLayer*> last_block;
  // encoder
  for (int l = 0; l < 4; ++l) {
    const int src = l == 0 ? in0 : pooled[l];
    const int cin = l == 0 ? cfg.spatial_channels : F[l - 1];
    const TRef skip{cat[l], 0, F[l]};
    const TRef xout = shared ? TRef{enc1[l], 0, F[l]} : skip;
    if (l > 0) MAU_TRY(add_pool(shared ? TRef{enc1[l - 1], 0, F[l - 1]} : TRef{cat[l - 1], 0, F[l - 1]}, TRef{pooled[l], 0, F[l - 1]}));
    MAU_TRY(add_vgg("conv" + std::to_string(l) + "_0", enc[l], src, 1, &zero, &cin, F[l], xout, l > 0, nullptr));
    if (shared) MAU_TRY(add_bcast(xout, skip));
  }
  if (shared) {
    MAU_TRY(add_pool(TRef{enc1[3], 0, F[3]}, TRef{pooled[4], 0, F[3]}));
    MAU_TRY(add_bcast(TRef{pooled[4], 0, F[3]}, TRef{bott, 0, F[3]}));
  } else {
    MAU_TRY(add_pool(TRef{cat[3], 0, F[3]}, TRef{bott, 0, F[3]}));
  }
  {
    TRef tdst{bott, F[3], cfg.temporal_dim}, mdst{bott, F[3] + (te ? cfg.temporal_dim : 0), cfg.meta_dim};
    MAU_TRY(build_embed_broadcast(&tdst, te ? 1 : 0, &mdst, me ? 1 : 0));     // joins the side stream
    const int cin = F[3] + E;
    MAU_TRY(add_vgg("conv4_0", enc[4], bott, 1, &zero, &cin, F[4], TRef{x4, 0, F[4]}, true, nullptr));
  }
  // decoder
  for (int l = 3; l >= 0; --l) {
    const TRef lower = l == 3 ? TRef{x4, 0, F[4]} : TRef{xdec[l + 1], 0, F[l + 1]};
    MAU_TRY(add_up(lower, TRef{cat[l], F[l], F[l + 1]}));
    const int cin = F[l] + F[l + 1];
    MAU_TRY(add_vgg("conv" + std::to_string(l) + "_1", dec[l], cat[l], 1, &zero, &cin, F[l], TRef{xdec[l], 0, F[l]}, true,
                    l == 0 ? &last_block : nullptr));
  }
  // head
  const TRef hx{xdec[0], 0, F[0]};
  note_op("head", &hx, nullptr, "\"w\":\"model.final.weight\",\"b\":\"model.final.bias\",\"index\":0");
  const int OC = cfg.out_channels;
  fwd_flops += 2.0 * OC * F[0] * (double)Hs[0] * Ws[0] * B;
  const bool fuse_head = !cfg.training && use_tc && conv_mode == MODE_HALO && OC <= 4 && F[0] <= 128 && !dry;
  if (fuse_head) {
    ConvLayer* L = last_block[1];
    L->head_w = fw; L->head_b = fb; L->head_oc = OC; L->head_tanh = OC == 2;
  } else if (!dry) {
    float* out_save = cfg.training ? static_cast<float*>(alloc(sizeof(float) * (size_t)B * OC * Hs[0] * Ws[0])) : nullptr;
    Op op; op.name = "head";
    const View vx = view(hx);
    const size_t obytes = sizeof(float) * (size_t)B * OC * Hs[0] * Ws[0];
    op.run = [=](Ctx& c) -> int {
      MAU_TRY(op_head(dt, vx, c.f(fw), c.f(fb), OC, OC == 2, c.out, c.st));
      if (out_save) MAU_CUDA(cudaMemcpyAsync(out_save, c.out, obytes, cudaMemcpyDeviceToDevice, c.st));
      return 0;
    };
    fwd.push_back(op);
    if (cfg.training)
      bwd_makers.push_back([=]() -> int {
        View gx;
        MAU_TRY(gview(hx, &gx));
        int acc = 0;
        MAU_TRY(gcontrib(hx, &acc));
        Op b; b.name = "head.bwd"; b.grad_first = fw; b.grad_last = fb;
        b.run = [=](Ctx& c) -> int {
          float* dw = c.g(fw); float* db = c.g(fb);
          if (!dw || !db) return fail("head backward needs final.weight / final.bias gradients");
          MAU_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * OC * F[0], c.st));
          MAU_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * OC, c.st));
          return op_head_bwd(dt, vx, c.f(fw), OC, OC == 2, out_save, c.gout, gx, dw, db, c.st);
        };
        bwd.push_back(b);
        return 0;
      });
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ U-Net++
int Plan::build_unetpp() {
  const int B = cfg.batch;
  const int* F = cfg.filters;
  const int E = cfg.temporal_dim + cfg.meta_dim;
  int Hs[5], Ws[5];
  Hs[0] = cfg.height; Ws[0] = cfg.width;
  for (int l = 1; l < 5; ++l) { Hs[l] = Hs[l - 1] / 2; Ws[l] = Ws[l - 1] / 2; }
  if (Hs[4] < 1 || Ws[4] < 1) return fail("tile too small for four 2x2 poolings");

  // state_dict order: reference src/model.py:63-96
  BlockIdx blk[5][5];
  for (int l = 0; l < 5; ++l)
    blk[l][0] = add_block_state("model.conv" + std::to_string(l) + "_0", l == 0 ? cfg.spatial_channels : F[l - 1], F[l], F[l]);
  for (int d = 1; d <= 4; ++d)
    for (int l = 0; l + d <= 4; ++l)
      blk[l][d] = add_block_state("model.conv" + std::to_string(l) + "_" + std::to_string(d), F[l] * d + F[l + 1] + E, F[l], F[l]);
  int lstm0, fc0, mlp0;
  add_encoder_state(&lstm0, &fc0, &mlp0);
  int fw[4], fb[4];
  const int nheads = cfg.deep_supervision ? 4 : 1;
  for (int i = 0; i < nheads; ++i) {
    const std::string nm = cfg.deep_supervision ? "model.final" + std::to_string(i + 1) : "model.final";
    fw[i] = add_state(nm + ".weight", (long long)cfg.out_channels * F[0], 0);
    fb[i] = add_state(nm + ".bias", cfg.out_channels, 0);
  }

  // level buffers: [x_l_0 .. x_l_{n-1} | up_1 .. up_n | emb],  n = 4 - l  (every torch.cat is a slice)
  const int Be = shared ? 1 : B;
  const int in0 = new_buf("maps_nhwc", Hs[0], Ws[0], cfg.spatial_channels, Be);
  int lv[4], nn[4], pooled[5], xlast[5], enc1[5] = {-1, -1, -1, -1, -1};
  if (shared)
    for (int l = 0; l < 5; ++l) enc1[l] = new_buf("x" + std::to_string(l) + "_0.shared", Hs[l], Ws[l], F[l], 1);
  for (int l = 0; l < 4; ++l) {
    nn[l] = 4 - l;
    lv[l] = new_buf("level" + std::to_string(l), Hs[l], Ws[l], nn[l] * F[l] + nn[l] * F[l + 1] + E);
    xlast[l] = new_buf("x" + std::to_string(l) + "_" + std::to_string(nn[l]), Hs[l], Ws[l], F[l]);
  }
  for (int l = 1; l < 5; ++l) pooled[l] = new_buf("pool" + std::to_string(l), Hs[l], Ws[l], F[l - 1], Be);
  xlast[4] = new_buf("x4_0", Hs[4], Ws[4], F[4]);
  auto xref = [&](int l, int j) -> TRef {      // where x_l_j lives
    if (l == 4) return TRef{xlast[4], 0, F[4]};
    if (j < nn[l]) return TRef{lv[l], j * F[l], F[l]};
    return TRef{xlast[l], 0, F[l]};
  };
  auto upref = [&](int l, int j) -> TRef { return TRef{lv[l], nn[l] * F[l] + (j - 1) * F[l + 1], F[l + 1]}; };
  auto embref = [&](int l) -> TRef { return TRef{lv[l], nn[l] * F[l] + nn[l] * F[l + 1], E}; };

  TRef tdst[4], mdst[4];
  for (int l = 0; l < 4; ++l) {
    tdst[l] = TRef{lv[l], embref(l).c0, cfg.temporal_dim};
    mdst[l] = TRef{lv[l], embref(l).c0 + cfg.temporal_dim, cfg.meta_dim};
  }
  MAU_TRY(build_encoders(lstm0, fc0, mlp0));
  { const TRef t{in0, 0, cfg.spatial_channels}; note_op("input", nullptr, &t); }
  if (!dry) {
    Op op; op.name = "nchw_to_nhwc";
    const View dst = whole(in0);
    op.run = [=](Ctx& c) -> int {
      if (c.maps_staged) {     // pre-staged NHWC tiles (mau_plan_forward_staged): a copy KERNEL instead of the layout kernel
        // (a cudaMemcpyAsync would queue on the copy engines behind the caller's own H2D prefetches: measured, e2e 7.4k vs 8.0k tiles/s)
        View src = dst;
        src.ptr = const_cast<void*>(c.maps_staged);
        View d8 = dst;
        src.C = d8.C = dst.cs;      // whole 8-channel groups, pad channel included
        return op_copy_slice(dt, src, d8, 0, c.st);
      }
      return op_nchw_to_nhwc(dt, c.maps, Be, cfg.spatial_channels, cfg.height, cfg.width, dst, c.st);
    };
    fwd.push_back(op);
  }
  auto add_bcast = [&](TRef src, TRef dst) -> int {
    note_op("bcast", &src, &dst);
    if (dry) return 0;
    Op op; op.name = "bcast." + bufs[src.buf].name;
    const View xs = view(src), yd = view(dst);
    op.run = [=](Ctx& c) -> int { return op_broadcast_batch(dt, xs, yd, c.st); };
    fwd.push_back(op);
    return 0;
  };
  auto add_pool = [&](TRef src, TRef dst) -> int {
    note_op("pool", &src, &dst);
    if (dry) return 0;
    Op op; op.name = "pool." + bufs[dst.buf].name;
    const View xs = view(src), yd = view(dst);
    op.run = [=](Ctx& c) -> int { return op_maxpool(dt, xs, yd, c.st); };
    fwd.push_back(op);
    if (cfg.training)
      bwd_makers.push_back([=]() -> int {
        View gy, gx;
        MAU_TRY(gview(dst, &gy));
        MAU_TRY(gview(src, &gx));
        int acc = 0;
        MAU_TRY(gcontrib(src, &acc));
        Op b; b.name = "pool.bwd." + bufs[dst.buf].name;
        b.run = [=](Ctx& c) -> int { return op_maxpool_bwd(dt, xs, gy, acc ? &gx : nullptr, gx, c.st); };
        bwd.push_back(b);
        return 0;
      });
    return 0;
  };
  // _upsample_match: single-stage resize to the exact target size (src/model.py:111-121)
  auto add_up = [&](TRef src, TRef dst) -> int {
    note_op("up", &src, &dst);
    BilinearTables t;
    MAU_TRY(make_bilinear(bufs[src.buf].H, bufs[src.buf].W, bufs[dst.buf].H, bufs[dst.buf].W, &t));
    if (dry) return 0;
    Op op; op.name = "up." + bufs[dst.buf].name + "@" + std::to_string(dst.c0);
    const View vs = view(src), vd = view(dst);
    op.run = [=](Ctx& c) -> int { return op_bilinear(dt, vs, vd, t, c.st); };
    fwd.push_back(op);
    if (cfg.training)
      bwd_makers.push_back([=]() -> int {
        View gd, gs;
        MAU_TRY(gview(dst, &gd));
        MAU_TRY(gview(src, &gs));
        int acc = 0;
        MAU_TRY(gcontrib(src, &acc));
        Op b; b.name = "up.bwd." + bufs[dst.buf].name + "@" + std::to_string(dst.c0);
        b.run = [=](Ctx& c) -> int { return op_bilinear_bwd(dt, gd, gs, t, acc, c.st); };
        bwd.push_back(b);
        return 0;
      });
    return 0;
  };
  const int zero = 0;
  auto encoder = [&](int l) -> int {
    const int src = l == 0 ? in0 : pooled[l];
    const int cin = l == 0 ? cfg.spatial_channels : F[l - 1];
    const TRef xout = shared ? TRef{enc1[l], 0, F[l]} : xref(l, 0);
    if (l > 0) MAU_TRY(add_pool(shared ? TRef{enc1[l - 1], 0, F[l - 1]} : xref(l - 1, 0), TRef{pooled[l], 0, F[l - 1]}));
    MAU_TRY(add_vgg("conv" + std::to_string(l) + "_0", blk[l][0], src, 1, &zero, &cin, F[l], xout, l > 0, nullptr));
    if (shared) MAU_TRY(add_bcast(xout, xref(l, 0)));
    return 0;
  };
  std::vector<ConvLayer*> last_block;
  auto node = [&](int l, int j) -> int {     // x_l_j, j >= 1
    MAU_TRY(add_up(xref(l + 1, j - 1), upref(l, j)));
    const int ss[3] = {0, upref(l, j).c0, embref(l).c0};
    const int sl[3] = {j * F[l], F[l + 1], E};
    if (cfg.training && !(cfg.flags & MAU_FLAG_EMB_DENSE_BWD)) next_emb_seg = 2;   // constant planes: closed-form backward
    return add_vgg("conv" + std::to_string(l) + "_" + std::to_string(j), blk[l][j], lv[l], 3, ss, sl, F[l], xref(l, j),
                   true, (l == 0 && j == 4) ? &last_block : nullptr);
  };
  // reference evaluation order, src/model.py:129-177
  MAU_TRY(encoder(0)); MAU_TRY(encoder(1));
  MAU_TRY(build_embed_broadcast(tdst, 4, mdst, 4));       // first consumer of the embeddings is x0_1: join here
  MAU_TRY(node(0, 1));
  MAU_TRY(encoder(2)); MAU_TRY(node(1, 1)); MAU_TRY(node(0, 2));
  MAU_TRY(encoder(3)); MAU_TRY(node(2, 1)); MAU_TRY(node(1, 2)); MAU_TRY(node(0, 3));
  MAU_TRY(encoder(4)); MAU_TRY(node(3, 1)); MAU_TRY(node(2, 2)); MAU_TRY(node(1, 3)); MAU_TRY(node(0, 4));

  const int OC = cfg.out_channels;
  const size_t osz = (size_t)B * OC * Hs[0] * Ws[0];
  fwd_flops += nheads * 2.0 * OC * F[0] * (double)Hs[0] * Ws[0] * B;
  for (int i = 0; i < nheads; ++i) {
    const TRef hx = cfg.deep_supervision ? xref(0, i + 1) : xref(0, 4);
    note_op("head", &hx, nullptr, "\"w\":\"" + state[fw[i]].name + "\",\"b\":\"" + state[fb[i]].name + "\",\"index\":" + std::to_string(i));
  }
  const bool fuse_head = !cfg.training && !cfg.deep_supervision && use_tc && conv_mode == MODE_HALO && OC <= 4 &&
                         F[0] <= 128 && !dry;
  if (fuse_head) {
    ConvLayer* L = last_block[1];
    L->head_w = fw[0]; L->head_b = fb[0]; L->head_oc = OC; L->head_tanh = OC == 2;
  } else if (!dry) {
    float* out_save = cfg.training ? static_cast<float*>(alloc(sizeof(float) * osz * nheads)) : nullptr;
    for (int i = 0; i < nheads; ++i) {
      const TRef hx = cfg.deep_supervision ? xref(0, i + 1) : xref(0, 4);
      const bool tanh0 = !cfg.deep_supervision && OC == 2;      // no activation on deep-supervision heads (:180-185)
      const View vx = view(hx);
      const int wi = fw[i], bi = fb[i];
      Op op; op.name = "head" + std::to_string(i);
      op.run = [=](Ctx& c) -> int {
        MAU_TRY(op_head(dt, vx, c.f(wi), c.f(bi), OC, tanh0, c.out + i * osz, c.st));
        if (out_save)
          MAU_CUDA(cudaMemcpyAsync(out_save + i * osz, c.out + i * osz, sizeof(float) * osz, cudaMemcpyDeviceToDevice, c.st));
        return 0;
      };
      fwd.push_back(op);
      if (cfg.training)
        bwd_makers.push_back([=]() -> int {
          View gx;
          MAU_TRY(gview(hx, &gx));
          int acc = 0;
          MAU_TRY(gcontrib(hx, &acc));
          if (acc) return fail("deep-supervision backward with shared head inputs is not supported");
          Op b; b.name = "head.bwd" + std::to_string(i); b.grad_first = wi; b.grad_last = bi;
          b.run = [=](Ctx& c) -> int {
            float* dw = c.g(wi); float* db = c.g(bi);
            if (!dw || !db) return fail("head backward needs final weight / bias gradients");
            MAU_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * OC * F[0], c.st));
            MAU_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * OC, c.st));
            return op_head_bwd(dt, vx, c.f(wi), OC, tanh0, out_save + i * osz, c.gout + i * osz, gx, dw, db, c.st);
          };
          bwd.push_back(b);
          return 0;
        });
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ build / run
int Plan::build() {
  if (cfg.model_type != MAU_MODEL_UNET && cfg.model_type != MAU_MODEL_UNETPP)
    return fail("Unsupported model_type: %d", cfg.model_type);
  if (cfg.batch < 1 || cfg.height < 16 || cfg.width < 16) return fail("need batch >= 1 and tiles >= 16x16");
  if (cfg.out_channels < 1 || cfg.out_channels > 8) return fail("out_channels must be in 1..8");
  for (int i = 0; i < 5; ++i)
    if (cfg.filters[i] < 8 || cfg.filters[i] % 8) return fail("filters must be positive multiples of 8");
  if (cfg.model_type == MAU_MODEL_UNETPP) { cfg.temporal_embeddings = 1; cfg.metadata_embeddings = 1; }
  else cfg.deep_supervision = 0;
  if ((cfg.temporal_embeddings && cfg.temporal_dim % 8) || (cfg.metadata_embeddings && cfg.meta_dim % 8))
    return fail("temporal_dim / meta_dim must be multiples of 8 (got %d / %d)", cfg.temporal_dim, cfg.meta_dim);
  { const int g = cfg.filters[0] / 8; if (g & (g - 1)) return fail("filters[0] must be 8 * 2^k for the 1x1 head"); }
  dt = cfg.precision == MAU_PRECISION_FP32 ? DT_F32 : DT_BF16;
  use_tc = (dt == DT_BF16) && !(cfg.flags & MAU_FLAG_CONV_FFMA);
  conv_mode = (cfg.flags & MAU_FLAG_CONV_TAPLOAD) ? MODE_TAP : ((cfg.flags & MAU_FLAG_CONV_ROW3) ? MODE_ROW3 : MODE_HALO);
  if (cfg.training && cfg.deep_supervision) return fail("deep supervision is forward-only in this engine");
  shared = (cfg.flags & MAU_FLAG_SHARED_MAPS) != 0;
  if (shared && cfg.training) return fail("MAU_FLAG_SHARED_MAPS is an inference-only fast path (BatchNorm batch statistics couple the rows in training)");
  // the gradient all-reduce only runs during backward: forward launches may use every SM, backward launches are
  // sized for (SMs - reserve)
  set_sm_reserve_override(0);
  int rc = cfg.model_type == MAU_MODEL_UNET ? build_unet() : build_unetpp();
  set_sm_reserve_override(-1);
  if (rc) return rc;
  if (cfg.training) {
    // the per-channel sums of every BatchNorm (forward statistics, then reused for the backward sums) live in ONE block that
    // is zeroed once per pass: 36 tiny memsets per step were 36 stream operations on the critical chain
    sums_bytes = 0;
    for (const ConvLayer* L : layers) sums_bytes += sizeof(double) * 2 * L->Cout;
    sums_all = static_cast<double*>(alloc(sums_bytes));
    if (!dry && !sums_all) return -1;
    {
      double* p = sums_all;
      for (ConvLayer* L : layers) { L->sums = dry ? nullptr : p; p += 2 * L->Cout; }
    }
    overlap_wgrad = !(cfg.flags & MAU_FLAG_NO_WGRAD_OVERLAP);
    conv_stats = !(cfg.flags & MAU_FLAG_NO_CONV_STATS);
    if (use_tc && !(cfg.flags & MAU_FLAG_WGRAD_V1)) {
      // one fp32 workspace [9][Cout][Cin] shared by all layers (backward ops are serialised on one stream)
      size_t fl = 0;
      for (const ConvLayer* L : layers) fl = std::max(fl, (size_t)9 * round_up(L->Cout, 4) * round_up(L->Cin, 4));
      wgrad_ws = static_cast<float*>(alloc(sizeof(float) * fl));
      if (dry) wgrad_ws = reinterpret_cast<float*>(16);
      else if (!wgrad_ws) return -1;
    }
    for (auto it = bwd_makers.rbegin(); it != bwd_makers.rend(); ++it) MAU_TRY((*it)());
    if (!dry) {
      // counters (num_batches_tracked) are bumped by one tiny kernel per forward
      counters_dev = static_cast<long long**>(alloc(sizeof(long long*) * std::max<size_t>(1, counter_idx.size())));
      if (!counters_dev) return -1;
    }
  }
  bwd_makers.clear();
  // describe
  std::ostringstream js;
  js.precision(15);
  js << "{\"model_type\":" << cfg.model_type << ",\"batch\":" << cfg.batch << ",\"height\":" << cfg.height
     << ",\"width\":" << cfg.width << ",\"training\":" << cfg.training << ",\"fwd_flops\":" << fwd_flops
     << ",\"exec_conv_flops\":" << exec_flops << ",\"workspace_bytes\":" << ws_bytes << ",\"state\":[";
  for (size_t i = 0; i < state.size(); ++i)
    js << (i ? "," : "") << "[\"" << state[i].name << "\"," << state[i].numel << "," << state[i].role << "]";
  js << "],\"layers\":[";
  for (size_t i = 0; i < layers.size(); ++i) {
    const ConvLayer* L = layers[i];
    js << (i ? "," : "") << "{\"name\":\"" << L->name << "\",\"cin\":" << L->Cin << ",\"cout\":" << L->Cout
       << ",\"h\":" << L->H << ",\"w\":" << L->W << ",\"kp\":" << L->Kp << ",\"flops\":" << L->flops << ",\"b\":" << L->B
       << ",\"segs\":[";
    for (int s = 0; s < L->nseg; ++s) js << (s ? "," : "") << "[" << L->seg_start[s] << "," << L->seg_len[s] << "]";
    js << "],\"in\":\"" << bufs[L->in_buf].name << "\",\"out\":\"" << bufs[L->out.buf].name << "\",\"out_c0\":" << L->out.c0
       << ",\"z\":\"" << (L->zbuf >= 0 ? bufs[L->zbuf].name : std::string()) << "\",\"emb_seg\":" << L->emb_seg
       << ",\"input_needs_grad\":" << (L->input_needs_grad ? 1 : 0) << ",\"weight\":\"" << state[L->iw].name << "\"}";
  }
  js << "],\"ops\":[";
  for (size_t i = 0; i < op_desc.size(); ++i) js << (i ? "," : "") << op_desc[i];
  js << "],\"buffers\":[";
  for (size_t i = 0; i < bufs.size(); ++i)
    js << (i ? "," : "") << "{\"name\":\"" << bufs[i].name << "\",\"h\":" << bufs[i].H << ",\"w\":" << bufs[i].W
       << ",\"c\":" << bufs[i].C << ",\"cs\":" << bufs[i].cs << ",\"b\":" << bufs[i].B << "}";
  js << "]}";
  describe_json = js.str();
  return 0;
}

static int run_ops(Plan* P, std::vector<Op>& ops, Ctx& c, bool backward) {
  std::vector<cudaEvent_t> ev;
  if (P->profiling) {
    ev.resize(ops.size() + 1);
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], c.st);
  }
  for (size_t i = 0; i < ops.size(); ++i) {
    int rc = ops[i].run(c);
    if (rc) return rc;
    if (P->profiling) cudaEventRecord(ev[i + 1], c.st);
    if (backward && P->hook && ops[i].grad_first >= 0) P->hook(P->hook_user, ops[i].grad_first, ops[i].grad_last);
  }
  if (P->profiling) {
    cudaEventSynchronize(ev.back());
    if (!backward) P->prof.clear();
    for (size_t i = 0; i < ops.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      P->prof.push_back({ops[i].name, ms});
    }
    for (auto& e : ev) cudaEventDestroy(e);
    for (auto& t : P->ktimers) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, t.e0, t.e1);
      P->prof.push_back({t.name, ms});
      cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
    }
    P->ktimers.clear();
  }
  return 0;
}

int Plan::run_forward(Ctx& c) {
  skip_pack = !cfg.training && state_version != 0 && state_version == packed_version && packed_state == last_state;
  packs_ahead = false;
  if (cfg.training) MAU_CUDA(cudaMemsetAsync(sums_all, 0, sums_bytes, c.st));
  if (cfg.training && overlap_wgrad && !profiling) {
    // training re-packs every weight tensor every step (the optimizer has just changed them): all 18 / 30 pack launches go
    // to the second stream up front and hide behind the first convolutions; each convolution waits for its own event
    cudaStream_t ws;
    MAU_TRY(w_fork(c, &ws));
    for (ConvLayer* L : layers) {
      if (!L->ev_pack) MAU_CUDA(cudaEventCreateWithFlags(&L->ev_pack, cudaEventDisableTiming));
      if (use_tc) MAU_TRY(conv_tc_pack_fwd(c.f(L->iw), L->Cout, L->Cin, L->kmap, L->Kp, L->wpack, ws));
      else MAU_TRY(conv_ffma_pack_fwd(c.f(L->iw), L->Cout, L->Cin, L->kmap, L->Kp, static_cast<float*>(L->wpack), ws));
      MAU_CUDA(cudaEventRecord(L->ev_pack, ws));
    }
    packs_ahead = true;
  }
  MAU_TRY(run_ops(this, fwd, c, false));
  MAU_TRY(side_join(c));
  if (!cfg.training) { packed_version = state_version; packed_state = last_state; }
  if (cfg.training && !counter_idx.empty()) {
    std::vector<long long*> ptrs;
    for (int idx : counter_idx) ptrs.push_back(static_cast<long long*>(c.state[idx]));
    // pageable-host async copies are staged by the runtime before returning
    MAU_CUDA(cudaMemcpyAsync(counters_dev, ptrs.data(), sizeof(long long*) * ptrs.size(), cudaMemcpyHostToDevice, c.st));
    MAU_TRY(op_bump_counters(counters_dev, (int)ptrs.size(), c.st));
  }
  forward_done = true;
  return 0;
}

int Plan::run_backward(Ctx& c) {
  if (!cfg.training) return fail("backward requires a training-mode plan");
  if (!forward_done) return fail("backward called before forward");
  if (emb_direct) MAU_CUDA(cudaMemsetAsync(demb, 0, sizeof(float) * cfg.batch * emb_dim, c.st));
  MAU_CUDA(cudaMemsetAsync(sums_all, 0, sums_bytes, c.st));
  MAU_TRY(run_ops(this, bwd, c, true));
  MAU_TRY(side_join(c));
  MAU_TRY(w_join(c.st));
  w_pending = false;
  forward_done = false;
  return 0;
}

}  // namespace mau
