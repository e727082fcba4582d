"""CPU: the oracle restatement (oracle/unet_oracle.py) against the golden vectors written
by the real reference (oracle/gen_golden.py) -- this is what pins the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import mau_b200
from oracle import unet_oracle as O

VARIANTS = {
    "unet_noemb": ("unet", 828, dict(temporal_embeddings=False, metadata_embeddings=False)),
    "unet_metaemb": ("unet", 828, dict(temporal_embeddings=False, metadata_embeddings=True)),
    "unet_emb": ("unet", 60, dict(temporal_embeddings=True, metadata_embeddings=True)),
    "unetpp_emb": ("unet++", 60, dict()),
}


def _small(name):
    mt, _, kw = VARIANTS[name]
    torch.manual_seed(123)
    m = mau_b200.UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 23, 37, 45, generator=g)
    ts = torch.randn(3, 40, generator=g)
    md = torch.randn(3, 8, generator=g)
    return mt, kw, m, x, ts, md


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_small_eval_matches_reference(name, golden_dir):
    mt, kw, m, x, ts, md = _small(name)
    want = np.load(os.path.join(golden_dir, f"small_{name}.npz"))["y_eval"]
    got = O.forward(m.state_dict(), mt, x, ts, md, training=False, **kw).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


def test_oracle_deep_supervision(golden_dir):
    torch.manual_seed(123)
    m = mau_b200.UrbanPredictor("unet++", 23, 828, 16, 8, 8, 32, 2, base_filters=8, deep_supervision=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 23, 37, 45, generator=g); ts = torch.randn(3, 40, generator=g); md = torch.randn(3, 8, generator=g)
    want = np.load(os.path.join(golden_dir, "small_unetpp_ds.npz"))["y_eval"]
    got = O.forward(m.state_dict(), "unet++", x, ts, md, deep_supervision=True)
    np.testing.assert_allclose(np.stack([t.numpy() for t in got]), want, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["unet_metaemb", "unetpp_emb"])
def test_oracle_full_width_kat(name, golden_dir):
    """Full 64-filter model at 50x50 (two-stage 12->24->25 resize), eval + one train step."""
    mt, T, kw = VARIANTS[name]
    kat = np.load(os.path.join(golden_dir, f"kat_{name}.npz"))
    torch.manual_seed(42)
    m = mau_b200.UrbanPredictor(mt, 23, 828, 64, 8, 64, 96, 2, **kw)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 23, 50, 50, generator=g); ts = torch.randn(2, T, generator=g); md = torch.randn(2, 8, generator=g)
    sd = m.state_dict()
    y = O.forward(sd, mt, x, ts, md, training=False, **kw).numpy()
    np.testing.assert_allclose(y, kat["y_eval"], rtol=1e-4, atol=1e-5)
    out, loss, grads, new_stats = O.train_step_grads(sd, mt, x, ts, md, None, loss="abs_mean", **kw)
    np.testing.assert_allclose(out.numpy(), kat["y_train"], rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - float(kat["loss"])) < 1e-5
    np.testing.assert_allclose(new_stats["model.conv0_0.bn1.running_mean"].numpy(), kat["bn_rm"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(new_stats["model.conv0_0.bn1.running_var"].numpy(), kat["bn_rv"], rtol=1e-4, atol=1e-6)
    assert int(new_stats["model.conv0_0.bn1.num_batches_tracked"]) == int(kat["bn_count"]) == 1
    want_norms = json.loads(str(kat["grad_norms"]))
    for k, wn in want_norms.items():
        if wn is None:
            assert grads[k] is None, k          # flag-disabled encoders: grad stays None
        else:
            assert abs(float(grads[k].norm()) - wn) <= 1e-3 * max(wn, 1e-3) + 1e-6, k
    for k in kat.files:
        if k.startswith("g::"):
            np.testing.assert_allclose(grads[k[3:]].numpy(), kat[k], rtol=2e-3, atol=2e-5)


def test_oracle_loss_terms(golden_dir):
    z = np.load(os.path.join(golden_dir, "loss_terms.npz"))
    pred = torch.tensor(z["pred"], requires_grad=True)
    tgt = torch.tensor(z["tgt"])
    d = O.loss_l1_gradient(pred, tgt, 0.1)
    assert abs(float(d["pixel"]) - float(z["l1"])) < 1e-6
    assert abs(float(d["gradient"]) - float(z["grad"])) < 1e-6
    (g,) = torch.autograd.grad(d["total"], pred)
    np.testing.assert_allclose(g.numpy(), z["dpred_l1"], rtol=1e-5, atol=1e-8)
    d2 = O.loss_mse_gradient(pred, tgt, 0.1)
    assert abs(float(d2["total"]) - float(z["total_mse"])) < 1e-6


def test_oracle_dw_map_known_answers():
    """test/evaluate.py:212-217: channel 0 is multiplied by 0, so class 0 wins only when all
    of channels 1..8 are <= 0; ties resolve to the lowest index."""
    x = np.zeros((1, 23, 2, 3), np.float32)
    x[0, 0, 0, 0] = 1.0                  # class 0 one-hot -> all products 0 -> argmax 0
    x[0, 5, 0, 1] = 1.0                  # class 5
    x[0, 3, 0, 2] = 1.0; x[0, 6, 0, 2] = 0.5   # 3*1 == 6*0.5 -> tie -> lowest index 3
    x[0, 8, 1, 0] = 1.0; x[0, 2, 1, 0] = 1.0   # 8 beats 2
    x[0, 4, 1, 1] = -1.0                 # negative product loses to zeros -> 0
    dw, rows = O.eval_metrics(x, np.zeros((1, 2, 2, 3), np.float32), np.ones((1, 2, 2, 3), np.float32))
    assert dw.dtype == np.int64
    assert dw[0].tolist() == [[0, 5, 3], [8, 0, 0]]
    overall = [r for r in rows if r[2] == -1]
    assert all(abs(r[4] - 1.0) < 1e-7 and abs(r[5] - 1.0) < 1e-7 for r in overall)
    assert sorted({r[2] for r in rows}) == [-1, 0, 3, 5, 8]


def test_oracle_laplacian_variance_equals_scipy():
    """test/evaluate.py:241-242 calls scipy.ndimage.laplace directly; the restatement must reproduce it."""
    from scipy.ndimage import laplace
    rng = np.random.default_rng(0)
    for shape in ((2, 2, 17, 23), (1, 2, 1, 9), (1, 1, 5, 1), (1, 2, 50, 50)):
        pred = rng.standard_normal(shape).astype(np.float32)
        tgt = (rng.standard_normal(shape) * 3 + 1).astype(np.float32)
        for stats in ((None, None), (14.5, 7.25)):
            got = O.laplacian_variance(pred, tgt, *stats)
            p, g = pred.copy(), tgt.copy()
            if stats[0] is not None and shape[1] > 1:
                p[:, 1] = p[:, 1] * np.float32(stats[1]) + np.float32(stats[0])
                g[:, 1] = g[:, 1] * np.float32(stats[1]) + np.float32(stats[0])
            for i in range(shape[0]):
                for ch in range(shape[1]):
                    lp, lg = laplace(p[i, ch]), laplace(g[i, ch])
                    assert lp.dtype == np.float32
                    # element-wise identical Laplacian (fp32), variance within fp32 summation error of np.var
                    pad = np.pad(p[i, ch].astype(np.float64), 1, mode="edge")
                    mine = ((pad[:-2, 1:-1] + pad[2:, 1:-1] - 2 * pad[1:-1, 1:-1]).astype(np.float32)
                            + (pad[1:-1, :-2] + pad[1:-1, 2:] - 2 * pad[1:-1, 1:-1]).astype(np.float32))
                    assert np.array_equal(mine, lp)
                    for k, ref in ((0, lp), (1, lg)):
                        want = float(np.var(ref))
                        assert abs(got[i, ch, k] - want) <= 2e-5 * max(abs(want), 1e-6), (shape, stats, i, ch, k)


def test_ssim_core_arithmetic_equals_torch_autograd(tmp_path):
    """csrc/ssim_core.h (the header the CUDA kernels include) driven serially on the host (oracle/ssim_host.cpp)
    against the torch restatement of piq's published SSIM: loss value and analytic gradient."""
    import ctypes as C
    import subprocess
    from oracle import ssim_oracle as S
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / "libssim_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", os.path.join(root, "oracle", "ssim_host.cpp"), "-o", lib])
    L = C.CDLL(lib)
    L.ssim_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 2, 23, 31), (1, 2, 11, 11), (3, 2, 40, 17), (2, 3, 30, 30), (1, 2, 64, 250)):
        out = (torch.randn(shape, generator=g) * 0.7).requires_grad_(True)
        tgt = torch.randn(shape, generator=g) * 0.7
        tgt[:, 0] = tgt[:, 0].clamp(-1, 1)                   # NDVI targets live in [-1, 1]
        loss = S.ssim_loss(out, tgt)
        loss.backward()
        o, t = out.detach().contiguous().numpy(), tgt.contiguous().numpy()
        grad, lv = np.zeros_like(o), C.c_double()
        assert L.ssim_host(o.ctypes.data, t.ctypes.data, *shape, C.byref(lv), grad.ctypes.data) == 0
        gref = out.grad.numpy()
        assert abs(lv.value - float(loss.detach())) < 1e-6
        assert np.abs(grad - gref).max() < 2e-5 * np.abs(gref).max()
        if shape[1] > 2:
            assert not grad[:, 2:].any() and not gref[:, 2:].any()      # only channels 0 and 1 enter the SSIM term
    x = torch.rand(1, 2, 20, 20, generator=g)
    assert abs(float(S.ssim_loss(x, x))) < 1e-6                          # identical maps: SSIM = 1
    L.ssim_pool_factor.restype = C.c_int                                  # piq: f = max(1, round(min(H, W) / 256)), Python's half-to-even round
    for m in list(range(11, 1300, 7)) + [383, 384, 385, 512, 639, 640, 641, 896, 1152]:
        assert L.ssim_pool_factor(m, m + 3) == max(1, round(m / 256)), m
    assert L.ssim_host(o.ctypes.data, t.ctypes.data, 1, 1, 20, 20, C.byref(lv), None) == 1     # needs both channels


def test_ssim_kernel_source_under_cpu_emulation_equals_torch_autograd(tmp_path):
    """csrc/ssim.cu itself -- the __global__ functions, their grid / block indexing, partial-block guards, the
    shared-memory tree reduction with __syncthreads and the launch sequence of op_ssim_loss -- compiled for the CPU on
    oracle/cuda_emu.h (256 real threads per block, barriers) and compared with torch autograd of the restatement.
    Output buffers start as NaN: every element must have been written."""
    import ctypes as C
    import subprocess
    from oracle import ssim_oracle as S
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / "libssim_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", os.path.join(root, "oracle", "ssim_kernels_emu.cpp"), "-o", lib])
    L = C.CDLL(lib)
    L.emu_ssim_work_floats.restype = C.c_longlong
    L.emu_ssim_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 2, 23, 31), (1, 2, 11, 11), (3, 2, 40, 17), (2, 3, 30, 30)):      # 273 / 1 / 210 / 400 windows per plane
        out = (torch.randn(shape, generator=g) * 0.7).requires_grad_(True)
        tgt = torch.randn(shape, generator=g) * 0.7
        tgt[:, 0] = tgt[:, 0].clamp(-1, 1)
        loss = S.ssim_loss(out, tgt)
        loss.backward()
        o, t = out.detach().contiguous().numpy(), tgt.contiguous().numpy()
        n = L.emu_ssim_work_floats(shape[0], shape[2], shape[3])
        assert n == 3 * shape[0] * 2 * (shape[2] - 10) * (shape[3] - 10)
        work, acc = np.full(n, np.nan, np.float32), np.full(1, np.nan, np.float64)
        lv, grad = np.full(1, np.nan, np.float32), np.full_like(o, np.nan)
        assert L.emu_ssim_loss(o.ctypes.data, t.ctypes.data, *shape, lv.ctypes.data, grad.ctypes.data, work.ctypes.data, acc.ctypes.data) == 0
        gref = out.grad.numpy()
        assert not np.isnan(grad).any() and not np.isnan(work).any()
        assert abs(float(lv[0]) - float(loss.detach())) < 1e-6
        assert np.abs(grad - gref).max() < 2e-5 * np.abs(gref).max()
    # the GPU check script itself (tools/ssim_gpu_check.py -> mau_b200.losses -> engine autograd wrappers), dry-run on the emulated kernels
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "ssim_dry_run.py"), lib], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"] and len(d["cases"]) == 5 and all(c["ok"] for c in d["cases"]), d
    # argument checks of the op: one channel only, tile smaller than the window
    z = np.zeros((1, 2, 40, 40), np.float32)
    for shp in ((1, 1, 20, 20), (1, 2, 10, 30)):
        assert L.emu_ssim_loss(z.ctypes.data, z.ctypes.data, *shp, lv.ctypes.data, None, work.ctypes.data, acc.ctypes.data) != 0
    # piq average-pools tiles with min(H, W) >= 384 by f = round(min / 256) before the SSIM (the app's 512 x 512 tiles: f = 2);
    # the pooled path (pool -> SSIM on the pooled planes -> un-pool with the scaling's chain rule) is exercised on small
    # tiles by forcing the factor, odd extents included (avg_pool2d drops the last row / column)
    for f, shape in ((2, (2, 2, 45, 52)), (3, (1, 3, 40, 47))):
        L.emu_ssim_force_pool(f)
        try:
            out = (torch.randn(shape, generator=g) * 0.7).requires_grad_(True)
            tgt = torch.randn(shape, generator=g) * 0.7
            loss = S.ssim_loss(out, tgt, force_pool=f)
            loss.backward()
            o, t = out.detach().contiguous().numpy(), tgt.contiguous().numpy()
            n = L.emu_ssim_work_floats(shape[0], shape[2], shape[3])
            work, acc = np.full(n, np.nan, np.float32), np.full(1, np.nan, np.float64)
            lv, grad = np.full(1, np.nan, np.float32), np.full_like(o, np.nan)
            assert L.emu_ssim_loss(o.ctypes.data, t.ctypes.data, *shape, lv.ctypes.data, grad.ctypes.data, work.ctypes.data, acc.ctypes.data) == 0
            gref = out.grad.numpy()
            assert not np.isnan(grad).any()
            assert abs(float(lv[0]) - float(loss.detach())) < 1e-6
            assert np.abs(grad - gref).max() < 2e-5 * np.abs(gref).max()
        finally:
            L.emu_ssim_force_pool(0)



def test_bilinear_backward_kernel_under_cpu_emulation_equals_torch_autograd(tmp_path):
    """csrc/bilinear_bwd_lean.cuh (streaming backward of the align_corners up-sampling: strip edges, sliding row window,
    the 2 / 4 / 6-contribution variants, branch-free absent entries) with the tables of csrc/bilinear_tables.h, compiled
    for the CPU on oracle/cuda_emu.h and compared with fp32 torch autograd of F.interpolate (reference src/model.py:12-17).
    gx starts as NaN (or as the addend): every element must be written."""
    import ctypes as C
    import subprocess
    import torch.nn.functional as F
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / "libbbl_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", os.path.join(root, "oracle", "bilinear_bwd_emu.cpp"), "-o", lib])
    L = C.CDLL(lib)
    L.emu_bilinear_bwd_lean.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p, C.c_int, C.c_int]
    g = torch.Generator().manual_seed(4)
    #        B  Hin Win  C  Hout Wout acc strip cs  c0
    cases = [(1, 5, 7, 8, 10, 14, 0, 4, 8, 0), (2, 13, 17, 16, 26, 35, 1, 8, 16, 0), (1, 31, 40, 24, 62, 80, 0, 16, 24, 0),
             (1, 30, 30, 8, 31, 31, 0, 8, 8, 0), (1, 15, 15, 64, 30, 30, 1, 4, 96, 16), (1, 20, 33, 8, 20, 33, 0, 16, 8, 0),
             (1, 12, 50, 8, 37, 125, 0, 5, 8, 0), (1, 2, 3, 8, 4, 6, 0, 4, 8, 0), (1, 9, 11, 128, 18, 22, 0, 4, 128, 0),
             (2, 7, 9, 40, 15, 19, 0, 4, 48, 8), (1, 62, 62, 8, 125, 125, 0, 16, 8, 0), (1, 124, 124, 8, 125, 125, 1, 16, 8, 0),
             (1, 6, 20, 8, 12, 9, 0, 4, 8, 0)]       # last: rows up-sampled, columns down-sampled (columns without a contribution)
    fans = set()
    for (B, Hin, Win, Cn, Hout, Wout, acc, strip, cs, c0) in cases:
        gy_full = torch.randn(B, Hout, Wout, cs, generator=g)
        x = torch.zeros(B, Cn, Hin, Win, requires_grad=True)
        F.interpolate(x, size=(Hout, Wout), mode="bilinear", align_corners=True).backward(gy_full[..., c0:c0 + Cn].permute(0, 3, 1, 2).contiguous())
        want = x.grad.permute(0, 2, 3, 1)
        init = torch.randn(B, Hin, Win, Cn, generator=g) if acc else torch.full((B, Hin, Win, Cn), float("nan"))
        gx = init.clone()
        fan = L.emu_bilinear_bwd_lean(gy_full.data_ptr(), B, Hin, Win, Cn, Hout, Wout, cs, c0, gx.data_ptr(), acc, strip)
        assert 1 <= fan <= 6
        fans.add(fan)
        ref = want + init if acc else want
        assert not torch.isnan(gx).any()
        assert (gx - ref).abs().max() < 4e-6, (B, Hin, Win, Cn, Hout, Wout)
    assert min(fans) <= 2 and any(3 <= f <= 4 for f in fans) and max(fans) >= 5      # every unrolling of the kernel was exercised
    # shapes the kernel does not serve are reported as such (the product then takes the table-driven kernels)
    z = torch.zeros(1, 9, 9, 8)
    assert L.emu_bilinear_bwd_lean(z.data_ptr(), 1, 2, 2, 8, 9, 9, 8, 0, z.data_ptr(), 0, 4) == 0      # 9 contributions per column
    assert L.emu_bilinear_bwd_lean(z.data_ptr(), 1, 9, 9, 8, 4, 4, 8, 0, z.data_ptr(), 0, 4) == 0      # rows down-sampled


def test_emulation_shim_reproduces_kernels_that_are_verified_on_hardware(tmp_path):
    """csrc/metrics.cu (class map, MAE / RMSE sums, Laplacian sums: all parity-green on a B200) under oracle/cuda_emu.h must
    agree with the oracle as well -- the control experiment for the emulated check of the SSIM kernels."""
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / "libmetrics_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-ffp-contract=off",
                           os.path.join(root, "oracle", "metrics_kernels_emu.cpp"), "-o", lib])
    L = C.CDLL(lib)
    L.emu_eval_metrics.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p]
    L.emu_laplacian_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    maps, _, _, tgt = O.synthetic_batch(2, 33, 29, T=4, seed=77)
    maps[0, 3, 0, 2] = 1.0; maps[0, 6, 0, 2] = 0.5; maps[0, 0:3, 0, 2] = 0; maps[0, 4:6, 0, 2] = 0; maps[0, 7:9, 0, 2] = 0   # tie 3*1 == 6*.5
    pred = torch.randn(2, 2, 33, 29, generator=torch.Generator().manual_seed(2))
    m, p, t = maps.contiguous().numpy(), pred.contiguous().numpy(), tgt.contiguous().numpy()
    want_dw, rows = O.eval_metrics(m, p, t, temp_mean=14.5, temp_std=7.25)
    dw, sums = np.full((2, 33, 29), -1, np.int64), np.full((2, 2, 10, 3), np.nan)
    assert L.emu_eval_metrics(m.ctypes.data, 23, p.ctypes.data, t.ctypes.data, 2, 2, 33, 29, 14.5, 7.25, dw.ctypes.data, sums.ctypes.data) == 0
    assert np.array_equal(dw, want_dw)
    for (i, ch, k, n, mae, rmse) in rows:
        slot = 0 if k < 0 else 1 + k
        assert int(sums[i, ch, slot, 0]) == n
        assert abs(sums[i, ch, slot, 1] / n - mae) < 2e-5 * max(1.0, abs(mae))
        assert abs(np.sqrt(sums[i, ch, slot, 2] / n) - rmse) < 2e-5 * max(1.0, abs(rmse))
    lap = np.full((2, 2, 4), np.nan)
    assert L.emu_laplacian_sums(p.ctypes.data, t.ctypes.data, 2, 2, 33, 29, 14.5, 7.25, lap.ctypes.data) == 0
    n = 33 * 29
    var = lap[..., 1::2] / n - (lap[..., 0::2] / n) ** 2
    np.testing.assert_allclose(var, O.laplacian_variance(p, t, 14.5, 7.25), rtol=2e-5, atol=1e-9)
