"""GPU diagnostics: kernel-by-kernel parity against torch, each group in its own process
(a CUDA fault or a hang in one group must not take the others down).

    python tools/gpu_diag.py            # run all groups
    python tools/gpu_diag.py conv_tc    # one group in-process
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["elementwise", "conv_ffma", "conv_tap", "conv_row3", "conv_halo", "wgrad", "lstm",
          "model_eval_fp32", "model_eval_ffma", "model_eval_tap", "model_eval", "model_train_ffma", "model_train",
          "model_full"]


def rel_err(a, b):
    import torch
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def nhwc(x, dtype, cs=None):
    """NCHW fp32 torch tensor -> NHWC tensor with channel stride cs."""
    import torch
    B, Cc, H, W = x.shape
    cs = cs or Cc
    y = torch.zeros(B, H, W, cs, device=x.device, dtype=dtype)
    y[..., :Cc] = x.permute(0, 2, 3, 1).to(dtype)
    return y.contiguous()


def run_conv(impl, dt, x, w, scale, shift, relu, cs_in=None, cs_out=None):
    import torch
    from mau_b200 import engine
    L = engine.lib()
    dtype = torch.bfloat16 if dt == 0 else torch.float32
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    cs_in = cs_in or ((Cin + 7) // 8 * 8)
    cs_out = cs_out or Cout
    xh = nhwc(x, dtype, cs_in)
    y = torch.zeros(B, H, W, cs_out, device=x.device, dtype=dtype)
    rc = L.mau_op_conv3x3(impl, dt, xh.data_ptr(), B, H, W, Cin, cs_in, w.contiguous().data_ptr(),
                          scale.data_ptr() if scale is not None else None,
                          shift.data_ptr() if shift is not None else None, int(relu), Cout, y.data_ptr(), cs_out,
                          None)
    engine.check(rc, "conv3x3")
    torch.cuda.synchronize()
    return y[..., :Cout].permute(0, 3, 1, 2).float()


def conv_cases():
    return [  # B, H, W, Cin, Cout
        (1, 16, 16, 64, 64), (2, 20, 24, 64, 64), (2, 15, 15, 128, 128), (1, 31, 31, 192, 256),
        (2, 33, 47, 23, 64), (1, 25, 25, 8, 8), (1, 62, 62, 320, 64), (2, 16, 16, 576, 512), (1, 50, 50, 64, 16),
        (3, 15, 15, 64, 192), (16, 31, 31, 128, 64), (5, 125, 125, 64, 128), (2, 250, 250, 64, 64),
    ]


def g_conv(impl, dt, tol):
    import torch
    import torch.nn.functional as F
    torch.manual_seed(0)
    worst = 0.0
    for (B, H, W, Cin, Cout) in conv_cases():
        x = torch.randn(B, Cin, H, W, device="cuda")
        w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)
        scale = torch.rand(Cout, device="cuda") + 0.5
        shift = torch.randn(Cout, device="cuda") * 0.1
        if dt == 0:
            xr, wr = x.bfloat16().float(), w.bfloat16().float()
        else:
            xr, wr = x, w
        ref = F.relu(F.conv2d(xr.double(), wr.double(), padding=1).float() * scale[None, :, None, None] + shift[None, :, None, None])
        t0 = time.time()
        got = run_conv(impl, dt, x, w, scale, shift, True)
        e = rel_err(got, ref)
        worst = max(worst, e)
        print(f"  conv impl={impl} dt={dt} B{B} {H}x{W} {Cin}->{Cout}: rel_err={e:.3e} ({time.time()-t0:.2f}s)", flush=True)
    print(f"  worst {worst:.3e} (tol {tol})")
    return worst < tol


def g_elementwise():
    import torch
    import torch.nn.functional as F
    from mau_b200 import engine
    L = engine.lib()
    ok = True
    torch.manual_seed(1)
    for dt, dtype, tol in ((1, torch.float32, 1e-6), (0, torch.bfloat16, 1e-2)):
        x = torch.randn(2, 24, 21, 35, device="cuda")
        xh = torch.zeros(2, 21, 35, 24, device="cuda", dtype=dtype)
        engine.check(L.mau_op_nchw_to_nhwc(dt, x.data_ptr(), 2, 24, 21, 35, 24, xh.data_ptr(), None))
        back = torch.zeros_like(x)
        engine.check(L.mau_op_nhwc_to_nchw(dt, xh.data_ptr(), 2, 24, 21, 35, 24, back.data_ptr(), None))
        e1 = rel_err(xh.float(), x.permute(0, 2, 3, 1)); e2 = rel_err(back, x)
        xq = xh.float().permute(0, 3, 1, 2)
        yp = torch.zeros(2, 10, 17, 24, device="cuda", dtype=dtype)
        engine.check(L.mau_op_maxpool2x2(dt, xh.data_ptr(), 2, 21, 35, 24, yp.data_ptr(), None))
        e3 = rel_err(yp.float().permute(0, 3, 1, 2), F.max_pool2d(xq, 2, 2))
        errs = [e1, e2, e3]
        for (ho, wo) in ((42, 70), (43, 71), (22, 35), (50, 50)):
            yb = torch.zeros(2, ho, wo, 24, device="cuda", dtype=dtype)
            engine.check(L.mau_op_bilinear(dt, xh.data_ptr(), 2, 21, 35, 24, ho, wo, yb.data_ptr(), None))
            errs.append(rel_err(yb.float().permute(0, 3, 1, 2), F.interpolate(xq, size=(ho, wo), mode="bilinear", align_corners=True)))
        torch.cuda.synchronize()
        print(f"  dt={dt}: to_nhwc {errs[0]:.2e} roundtrip {errs[1]:.2e} pool {errs[2]:.2e} bilinear {[f'{e:.2e}' for e in errs[3:]]}")
        ok &= all(e <= tol for e in errs)
    return ok


def g_wgrad():
    import torch
    import torch.nn.functional as F
    from mau_b200 import engine
    L = engine.lib()
    torch.manual_seed(2)
    ok = True
    for impl, dt in ((2, 1), (2, 0), (0, 0)):
        dtype = torch.bfloat16 if dt == 0 else torch.float32
        for (B, H, W, Cin, Cout) in [(1, 16, 16, 64, 128), (2, 20, 24, 64, 64), (2, 15, 15, 128, 256), (2, 33, 47, 24, 64), (1, 25, 25, 8, 8)]:
            x = torch.randn(B, Cin, H, W, device="cuda"); dy = torch.randn(B, Cout, H, W, device="cuda")
            if dt == 0:
                x, dy = x.bfloat16().float(), dy.bfloat16().float()
            xd = x.double().requires_grad_(False)
            w = torch.zeros(Cout, Cin, 3, 3, device="cuda", dtype=torch.double, requires_grad=True)
            (F.conv2d(xd, w, padding=1) * dy.double()).sum().backward()
            ref = w.grad.float()
            dw = torch.zeros(Cout, Cin, 3, 3, device="cuda")
            xh, dyh = nhwc(x, dtype), nhwc(dy, dtype)
            engine.check(L.mau_op_conv3x3_wgrad(impl, dt, xh.data_ptr(), dyh.data_ptr(), B, H, W, Cin, Cin, Cout, Cout, dw.data_ptr(), None), "wgrad")
            torch.cuda.synchronize()
            e = rel_err(dw, ref)
            print(f"  wgrad impl={impl} dt={dt} B{B} {H}x{W} {Cin}->{Cout}: rel_err={e:.3e}", flush=True)
            ok &= e < (1e-4 if dt == 1 else 5e-3)
    return ok


def g_lstm():
    import torch
    from mau_b200 import engine
    L = engine.lib()
    ok = True
    for Hd, T, B in ((96, 828, 3), (32, 60, 2), (16, 5, 1)):
        torch.manual_seed(3)
        lstm = torch.nn.LSTM(1, Hd, batch_first=True).cuda()
        s = torch.randn(B, T, device="cuda")
        with torch.no_grad():
            _, (h, _) = lstm(s.unsqueeze(-1))
        out = torch.zeros(B, Hd, device="cuda")
        engine.check(L.mau_op_lstm_last_hidden(s.data_ptr(), B, T, Hd, lstm.weight_ih_l0.data_ptr(), lstm.weight_hh_l0.data_ptr(),
                                               lstm.bias_ih_l0.data_ptr(), lstm.bias_hh_l0.data_ptr(), out.data_ptr(), None), "lstm")
        torch.cuda.synchronize()
        e = rel_err(out, h[-1])
        print(f"  lstm Hd={Hd} T={T}: rel_err={e:.3e}")
        ok &= e < 1e-4
    return ok


def _model_case(variant, small, training, precision, flags=0, B=2, HW=(37, 45), T=40):
    import torch
    import mau_b200
    from oracle import unet_oracle as O
    mt, kw = variant
    torch.manual_seed(123)
    if small:
        args = (23, 828, 16, 8, 8, 32, 2); extra = dict(base_filters=8)
    else:
        args = (23, 828, 64, 8, 64, 96, 2); extra = {}
    m = mau_b200.UrbanPredictor(mt, *args, **extra, **kw)
    O.perturb_bn_stats(m.state_dict())
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 23, *HW, generator=g); ts = torch.randn(B, T, generator=g); md = torch.randn(B, 8, generator=g)
    sd_cpu = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    m.set_precision(precision)
    if flags:
        m.model._cfg["flags"] = flags
    xc, tc, mc = x.cuda(), ts.cuda(), md.cuda()
    if not training:
        m.eval()
        with torch.no_grad():
            y = m(xc, tc, mc)
        torch.cuda.synchronize()
        ref = O.forward(sd_cpu, mt, x, ts, md, training=False, **kw)
        return {"out": rel_err(y.cpu(), ref)}
    m.train()
    y = m(xc, tc, mc)
    tgt = torch.randn(y.shape, generator=g)
    loss = (y - tgt.cuda()).abs().mean()
    loss.backward()
    torch.cuda.synchronize()
    out, lv, grads, new_stats = O.train_step_grads(sd_cpu, mt, x, ts, md, tgt, loss="l1", **kw)
    res = {"out": rel_err(y.detach().cpu(), out), "loss": abs(float(loss) - float(lv)) / abs(float(lv))}
    worst, worst_name = 0.0, ""
    errs = []
    for n, p in m.named_parameters():
        gref = grads[n]
        if gref is None:
            assert p.grad is None, f"{n} should have no grad"
            continue
        assert p.grad is not None, f"{n} has no grad"
        e = float((p.grad.cpu() - gref).norm() / gref.norm().clamp_min(1e-6))
        if gref.norm() < 1e-5:   # conv biases in front of BN: analytically zero gradient
            e = float((p.grad.cpu() - gref).abs().max())
        errs.append((e, n))
        if e > worst:
            worst, worst_name = e, n
    errs.sort(reverse=True)
    res["grad_top5"] = [(round(e, 5), n) for e, n in errs[:12]]
    res["grad_median"] = errs[len(errs) // 2][0]
    res["grad_worst"] = worst
    res["grad_worst_name"] = worst_name
    sd = m.state_dict()
    res["running_mean"] = rel_err(sd["model.conv0_0.bn1.running_mean"].cpu(), new_stats["model.conv0_0.bn1.running_mean"])
    res["running_var"] = rel_err(sd["model.conv1_0.bn2.running_var"].cpu(), new_stats["model.conv1_0.bn2.running_var"])
    res["nbt"] = int(sd["model.conv0_0.bn1.num_batches_tracked"])
    return res


VARS = {
    "unet_noemb": ("unet", dict(temporal_embeddings=False, metadata_embeddings=False)),
    "unet_meta": ("unet", dict(temporal_embeddings=False, metadata_embeddings=True)),
    "unet_emb": ("unet", dict(temporal_embeddings=True, metadata_embeddings=True)),
    "unetpp": ("unet++", dict()),
}


def g_model(training, precision, flags, tol, names=("unet_noemb", "unet_meta", "unet_emb", "unetpp"), small=True, **kw):
    ok = True
    for n in names:
        try:
            r = _model_case(VARS[n], small, training, precision, flags, **kw)
        except Exception as e:  # noqa
            print(f"  {n}: EXCEPTION {type(e).__name__}: {e}", flush=True)
            ok = False
            continue
        print(f"  {n}: {r}", flush=True)
        ok &= r["out"] < tol and r.get("grad_worst", 0) < max(tol * 5, 5e-2 if precision == "bf16" else 1e-3)
    return ok


def run_group(name):
    if name == "elementwise": return g_elementwise()
    if name == "conv_ffma": return g_conv(2, 1, 1e-5) and g_conv(2, 0, 1e-2)
    if name == "conv_tap": return g_conv(1, 0, 1e-2)
    if name == "conv_row3": return g_conv(3, 0, 1e-2)
    if name == "conv_halo": return g_conv(0, 0, 1e-2)
    if name == "wgrad": return g_wgrad()
    if name == "lstm": return g_lstm()
    if name == "model_eval_fp32": return g_model(False, "fp32", 0, 2e-5)
    if name == "model_eval_ffma": return g_model(False, "bf16", 4, 3e-2)
    if name == "model_eval_tap": return g_model(False, "bf16", 2, 3e-2)
    if name == "model_eval": return g_model(False, "bf16", 0, 3e-2)
    if name == "model_train_ffma": return g_model(True, "fp32", 0, 1e-4)
    if name == "model_train": return g_model(True, "bf16", 0, 3e-2)
    if name == "model_train_full_fp32":
        return g_model(True, "fp32", 0, 1e-4, names=("unet_noemb", "unet_meta", "unet_emb"), small=False, B=4, HW=(64, 64), T=60)
    if name == "model_train_full_fp32_b2":
        return g_model(True, "fp32", 0, 1e-4, names=("unet_meta",), small=False, B=2, HW=(50, 50), T=60)
    if name == "model_train_bf16_ffma": return g_model(True, "bf16", 4, 3e-2)
    if name == "model_train_full":
        return g_model(True, "bf16", 0, 3e-2, names=("unet_meta", "unetpp"), small=False, B=4, HW=(64, 64), T=60)
    if name == "model_train_full_ffma":
        return g_model(True, "bf16", 4, 3e-2, names=("unet_meta",), small=False, B=4, HW=(64, 64), T=60)
    if name == "model_eval_row3": return g_model(False, "bf16", 512, 3e-2)
    if name == "model_full":
        return g_model(False, "bf16", 0, 3e-2, names=("unet_meta", "unetpp"), small=False, B=2, HW=(50, 50), T=60)
    raise SystemExit(f"unknown group {name}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "--only":
        ok = run_group(sys.argv[1])
        print("GROUP", sys.argv[1], "PASS" if ok else "FAIL", flush=True)
        sys.exit(0 if ok else 1)
    groups = sys.argv[2].split(",") if len(sys.argv) > 2 else GROUPS
    summary = {}
    for g in groups:
        print(f"=== {g}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=240, capture_output=True, text=True)
            out = r.stdout + ("\nSTDERR: " + r.stderr[-1500:] if r.returncode else "")
            summary[g] = "PASS" if r.returncode == 0 else f"FAIL({r.returncode})"
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            summary[g] = "TIMEOUT"
        print(out, flush=True)
        print(f"--- {g}: {summary[g]} in {time.time()-t0:.1f}s", flush=True)
    print("SUMMARY", summary)
