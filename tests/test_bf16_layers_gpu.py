"""Teacher-forced, per-launch parity of a whole training step (forward + backward) in bf16 mode.

Why this test exists.  A whole-model comparison of bf16 training against ANY oracle cannot be tight: with batch-statistics
BatchNorm in front of every ReLU the per-tensor gradient of these networks moves by 20-30 % when 1e-4 of the stored bf16
values flip by one ulp (two CPU emulations of the same bf16 storage points that differ only in fp32 vs fp64 arithmetic
between the stores disagree by that much -- tools/bf16_sensitivity.py, profiles/r02_bf16_sensitivity.md), so a loose
whole-model bound would also pass with one wrong kernel.  Here every kernel launch of the step is checked IN SITU instead:
the plan's own activation / gradient buffers are read back (mau_plan_buffer_ptr) and each op is recomputed on the CPU in
fp32 from the *device's own inputs to that op* with bf16 rounding exactly where the engine stores bf16:

  forward   conv (+bias) -> z | batch-stat BN + ReLU -> y | max-pool | bilinear stages | embedding planes | 1x1 head + tanh
  backward  BN+ReLU backward -> dz, dgamma, dbeta | weight gradient from (x, dz) | data gradient from (dz, W) summed with the
            pool / bilinear / head contributions into every gradient buffer | embedding gradient -> MLP / LSTM parameters

One flipped bf16 ulp is 0.4-0.8 % of one element; the assertions are relative L2 per tensor (2e-3 for bf16 stores, 1e-3 for
fp32 outputs), which a wrong tap, segment offset, accumulation flag or layer wiring exceeds by orders of magnitude.
The same harness runs in fp32 mode with 1e-4 tolerances (a check of the harness itself).
Reference semantics cited: src/model.py:9-21 (VGGBlock), :98-121, :219, :243-259 (embeddings, resizes), :284-292 (head).
"""
import pytest
import torch
import torch.nn.functional as F

import mau_b200
from mau_b200 import engine
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def _nchw(t):        # [B,H,W,C] -> [B,C,H,W]
    return t.permute(0, 3, 1, 2).contiguous()


def _bn_name(conv_weight):
    """'model.conv1_0.conv2.weight' -> 'model.conv1_0.bn2' (VGGBlock, src/model.py:12-15)."""
    blk, cv, _ = conv_weight.rsplit(".", 2)
    return f"{blk}.bn{cv[-1]}"


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


class _Dump:
    """Host copies of the plan's buffers, sliced by (buffer name, first channel, channels) -> NCHW fp32."""

    def __init__(self, plan, grad):
        self.t = {}
        for b in plan.describe()["buffers"]:
            v = plan.buffer(b["name"], grad=grad)
            if v is not None:
                self.t[b["name"]] = v.float().cpu()

    def has(self, name):
        return name in self.t

    def get(self, name, c0, C):
        return _nchw(self.t[name][..., c0:c0 + C])


def _run_step(mt, kw, ctor, bf, B, H, W, T, precision, seed):
    torch.manual_seed(seed)
    m = mau_b200.UrbanPredictor(mt, *ctor, base_filters=bf, **kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x, ts, md, tgt = O.synthetic_batch(B, H, W, T=T, seed=1000 + seed)
    m = m.cuda().set_precision(precision).train()
    out = m(x.cuda(), ts.cuda(), md.cuda())
    out.retain_grad()
    plan = next(p for p in m.model._plans.values() if p.cfg["training"])
    fwd = _Dump(plan, grad=False)                         # z buffers hold z now (backward overwrites them with dz)
    loss = engine.compute_loss_mse_gradient(out, tgt.cuda(), 0.1)["total"]
    loss.backward()
    torch.cuda.synchronize()
    bwd_act = _Dump(plan, grad=False)                     # z buffers now hold dz
    bwd_g = _Dump(plan, grad=True)
    grads = {n: (p.grad.detach().cpu() if p.grad is not None else None) for n, p in m.named_parameters()}
    return dict(desc=plan.describe(), sd=sd, x=x, ts=ts, md=md, out=out.detach().cpu(), gout=out.grad.detach().cpu(),
                fwd=fwd, dz=bwd_act, g=bwd_g, grads=grads, mt=mt, kw=kw)


CASES = {
    # name: (model_type, ctor kwargs, ctor dims, base_filters, B, H, W, T)
    "unet_small_odd": ("unet", dict(temporal_embeddings=True, metadata_embeddings=True), (23, 828, 16, 8, 8, 32, 2), 8, 3, 37, 45, 24),
    "unet_full_width": ("unet", dict(temporal_embeddings=False, metadata_embeddings=True), (23, 828, 64, 8, 64, 96, 2), 64, 2, 64, 64, 8),
    "unetpp_small_odd": ("unet++", dict(), (23, 828, 16, 8, 8, 32, 2), 8, 2, 37, 45, 24),
    "unetpp_width32": ("unet++", dict(), (23, 828, 32, 8, 32, 32, 2), 32, 2, 48, 48, 16),
}


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("case", list(CASES))
def test_every_launch_of_a_training_step_teacher_forced(case, precision):
    mt, kw, ctor, bf, B, H, W, T = CASES[case]
    r = _run_step(mt, kw, ctor, bf, B, H, W, T, precision, seed=7)
    bf16 = precision == "bf16"
    rnd = O._bf16 if bf16 else (lambda t: t)
    tol_store = 2e-3 if bf16 else 1e-4        # a tensor the engine stores in the activation dtype
    tol_f32 = 1e-3 if bf16 else 1e-4          # an fp32 output accumulated from stored values (weight gradients, sums)
    desc, sd, fwd, dzd, gd, grads = r["desc"], r["sd"], r["fwd"], r["dz"], r["g"], r["grads"]
    bufs = {b["name"]: b for b in desc["buffers"]}
    fails = []

    def check(what, got, want, tol):
        e = _rel(got, want)
        if not e <= tol:
            fails.append(f"{what}: rel L2 {e:.3e} > {tol:.0e}")

    # ---- embeddings (fp32 encoders) ----------------------------------------------------------------------------
    te = mt == "unet++" or kw.get("temporal_embeddings", True)
    me = mt == "unet++" or kw.get("metadata_embeddings", True)
    enc_params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
                  if ("temporal_encoder" in k or "meta_encoder" in k) and v.is_floating_point()}
    emb_val = {}
    if te:
        emb_val["temporal"] = O.temporal_encoder(enc_params, "model.temporal_encoder", r["ts"])
    if me:
        emb_val["metadata"] = O.metadata_encoder(enc_params, "model.meta_encoder", r["md"])

    # ---- forward, op by op -------------------------------------------------------------------------------------
    contrib = {}      # gradient contributions per buffer, accumulated by the backward checks below

    def add_contrib(name, c0, t_nchw):
        b = bufs[name]
        if name not in contrib:
            contrib[name] = [torch.zeros(b["b"], b["c"], b["h"], b["w"]), torch.zeros(b["c"], dtype=torch.bool)]
        contrib[name][0][:, c0:c0 + t_nchw.shape[1]] += t_nchw
        contrib[name][1][c0:c0 + t_nchw.shape[1]] = True

    emb_dsts = []
    for op in desc["ops"]:
        k = op["kind"]
        if k == "input":
            n, c0, C = op["dst"]
            check("input layout", fwd.get(n, c0, C), rnd(r["x"]), 0.0 if bf16 else 1e-7)
        elif k == "pool":
            src, dst = fwd.get(*op["src"]), fwd.get(*op["dst"])
            if not torch.equal(F.max_pool2d(src, 2, 2), dst):
                fails.append(f"pool {op['src'][0]} -> {op['dst'][0]} is not bit-exact")
        elif k == "up":
            src, dst = fwd.get(*op["src"]), fwd.get(*op["dst"])
            want = rnd(F.interpolate(src, size=dst.shape[2:], mode="bilinear", align_corners=True))
            check(f"bilinear {op['src'][0]} -> {op['dst'][0]}@{op['dst'][1]}", dst, want, tol_store)
        elif k == "emb":
            n, c0, C = op["dst"]
            dst = fwd.get(n, c0, C)
            want = rnd(emb_val[op["which"]].detach())[:, :, None, None].expand_as(dst)
            check(f"embedding planes {op['which']} -> {n}@{c0}", dst, want, tol_store)
            emb_dsts.append((op["which"], n, c0, C))
        elif k == "head":
            src = fwd.get(*op["src"])
            want = O.head({"h.weight": sd[op["w"]], "h.bias": sd[op["b"]]}, "h", src)
            check("1x1 head + tanh", r["out"], want, 1e-4)

    for L in desc["layers"]:
        xin = torch.cat([fwd.get(L["in"], s0, n) for s0, n in L["segs"]], 1)
        w = rnd(sd[L["weight"]])
        bias = sd[L["weight"][:-len("weight")] + "bias"]
        z = fwd.get(L["z"], 0, L["cout"])
        check(f"{L['name']} conv -> z", z, rnd(F.conv2d(xin, w, bias, padding=1)), tol_store)
        bn = _bn_name(L["weight"])
        gam, bet = sd[bn + ".weight"], sd[bn + ".bias"]
        y_ref = F.relu(F.batch_norm(z, None, None, gam, bet, True, 0.1, 1e-5))
        check(f"{L['name']} BN+ReLU -> y", fwd.get(L["out"], L["out_c0"], L["cout"]), rnd(y_ref), tol_store)

    # ---- backward, consumer by consumer ------------------------------------------------------------------------
    for op in desc["ops"]:
        if op["kind"] == "head":
            src = fwd.get(*op["src"]).requires_grad_(True)
            wp = sd[op["w"]].clone().requires_grad_(True)
            bp = sd[op["b"]].clone().requires_grad_(True)
            O.head({"h.weight": wp, "h.bias": bp}, "h", src).backward(r["gout"])
            add_contrib(op["src"][0], op["src"][1], src.grad)
            check("head dW", grads[op["w"]], wp.grad, tol_f32)
            check("head db", grads[op["b"]], bp.grad, tol_f32)
    # the remaining ops consume gradient buffers the device filled; process consumers in reverse forward order so that
    # every contribution is recomputed from the DEVICE's own upstream gradient (teacher forcing)
    demb = {k: torch.zeros_like(v) for k, v in emb_val.items()}
    for L in reversed(desc["layers"]):
        cout = L["cout"]
        if not gd.has(L["out"]):
            fails.append(f"{L['name']}: no gradient buffer for {L['out']}")
            continue
        gy = gd.get(L["out"], L["out_c0"], cout).double()
        z = fwd.get(L["z"], 0, cout).double()
        bn = _bn_name(L["weight"])
        gam = sd[bn + ".weight"].double()
        # BatchNorm (batch statistics) + ReLU backward in closed form, fp64, with the ReLU mask taken from the DEVICE's own
        # activation (y > 0): an element whose pre-activation is within rounding of zero may legitimately fall on either
        # side, and one flipped element moves a whole channel's dbeta by ~1/sqrt(pixels) -- the mask is teacher-forced too
        mask = (fwd.get(L["out"], L["out_c0"], cout) > 0).double()
        mean = z.mean((0, 2, 3), keepdim=True)
        rstd = 1.0 / torch.sqrt(z.var((0, 2, 3), unbiased=False, keepdim=True) + 1e-5)
        xh = (z - mean) * rstd
        gt = gy * mask
        dbeta_ref, dgamma_ref = gt.sum((0, 2, 3)), (gt * xh).sum((0, 2, 3))
        dz_ref = gam[None, :, None, None] * rstd * (gt - gt.mean((0, 2, 3), keepdim=True) - xh * (gt * xh).mean((0, 2, 3), keepdim=True))
        dz_dev = dzd.get(L["z"], 0, cout)
        check(f"{L['name']} BN+ReLU backward -> dz", dz_dev, rnd(dz_ref.float()), tol_store)
        check(f"{L['name']} dgamma", grads[bn + ".weight"], dgamma_ref, tol_f32)
        check(f"{L['name']} dbeta", grads[bn + ".bias"], dbeta_ref, tol_f32)
        gb = grads[L["weight"][:-len("weight")] + "bias"]
        if float(gb.abs().max()) != 0.0:
            fails.append(f"{L['name']} conv bias gradient must be exactly 0 under batch-statistics BatchNorm")
        xin = torch.cat([fwd.get(L["in"], s0, n) for s0, n in L["segs"]], 1)
        w = rnd(sd[L["weight"]])
        dw_ref = torch.nn.grad.conv2d_weight(xin, w.shape, dz_dev, padding=1)
        dw_dev = grads[L["weight"]]
        ci = 0
        for si, (s0, n) in enumerate(L["segs"]):
            closed_form = si == L["emb_seg"]       # U-Net++ embedding planes: closed form from fp32 embeddings / weights
            check(f"{L['name']} dW[:, {ci}:{ci + n}]", dw_dev[:, ci:ci + n], dw_ref[:, ci:ci + n], 1e-2 if closed_form and bf16 else tol_f32)
            ci += n
        if L["input_needs_grad"]:
            dx = torch.nn.grad.conv2d_input(xin.shape, w, dz_dev, padding=1)
            ci = 0
            for si, (s0, n) in enumerate(L["segs"]):
                if si == L["emb_seg"]:     # closed form on the device: no gradient planes, the spatial sum goes to d emb
                    which_t = [(wh, c0, C) for wh, nme, c0, C in emb_dsts if nme == L["in"] and s0 <= c0 < s0 + n]
                    for wh, c0, C in which_t:
                        demb[wh] += dx[:, ci + c0 - s0: ci + c0 - s0 + C].sum((2, 3))
                else:
                    add_contrib(L["in"], s0, dx[:, ci:ci + n])
                ci += n
    for op in reversed(desc["ops"]):
        k = op["kind"]
        if k not in ("pool", "up"):
            continue
        if not gd.has(op["dst"][0]):
            continue
        g_dst = gd.get(*op["dst"])
        src = fwd.get(*op["src"]).requires_grad_(True)
        if k == "pool":
            F.max_pool2d(src, 2, 2).backward(g_dst)
        else:
            F.interpolate(src, size=g_dst.shape[2:], mode="bilinear", align_corners=True).backward(g_dst)
        add_contrib(op["src"][0], op["src"][1], src.grad)

    # every gradient buffer = the rounded sum of its consumers' contributions
    for name, (tot, covered) in contrib.items():
        if not gd.has(name):
            fails.append(f"gradient buffer {name} missing")
            continue
        b = bufs[name]
        dev = gd.get(name, 0, b["c"])
        # embedding-plane slices fed densely (U-Net bottleneck) are compared too; closed-form slices were never written
        idx = covered.nonzero().flatten()
        check(f"gradient buffer {name} ({len(idx)} ch)", dev[:, idx], rnd(tot[:, idx]), 2 * tol_store)

    # ---- embedding gradient -> encoder parameters ----------------------------------------------------------------
    for which, n, c0, C in emb_dsts:
        if (n in contrib) and bool(contrib[n][1][c0:c0 + C].all()) and gd.has(n):
            demb[which] += gd.get(n, c0, C).sum((2, 3))        # dense path: device gradient planes, summed
    if emb_val:
        tot = sum((emb_val[k] * demb[k]).sum() for k in emb_val)
        tot.backward()
        for k, p in enc_params.items():
            used = ("temporal_encoder" in k and te) or ("meta_encoder" in k and me)
            if not used:
                if grads[k] is not None:
                    fails.append(f"{k}: flag-disabled encoder must keep grad None")
                continue
            check(f"encoder gradient {k}", grads[k], p.grad, 2e-2 if bf16 else 1e-3)

    assert not fails, f"{len(fails)} mismatches:\n" + "\n".join(fails[:40])
