"""CPU: the drop-in boundary (module surface, state_dict layout), the C-ABI library's exports and
the host-side layer-graph logic (no kernel runs here)."""
import json
import os
import re

import pytest
import torch

import mau_b200
from mau_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTOR = (23, 828, 64, 8, 64, 96, 2)
VARIANTS = {
    "unet_noemb": ("unet", dict(temporal_embeddings=False, metadata_embeddings=False)),
    "unet_metaemb": ("unet", dict(temporal_embeddings=False, metadata_embeddings=True)),
    "unet_emb": ("unet", dict(temporal_embeddings=True, metadata_embeddings=True)),
    "unetpp_emb": ("unet++", dict()),
}


@pytest.fixture(scope="module")
def golden_keys(golden_dir):
    return json.load(open(os.path.join(golden_dir, "state_keys.json")))


@pytest.mark.parametrize("name", list(VARIANTS))
def test_state_dict_layout_matches_reference(name, golden_keys):
    mt, kw = VARIANTS[name]
    m = mau_b200.UrbanPredictor(mt, *CTOR, **kw)
    got = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert got == golden_keys[name]
    # optimizer state indexes parameters by registration order
    pk = [k for k, _ in m.named_parameters()]
    assert pk == [k for k, _, _ in golden_keys[name] if "running_" not in k and "num_batches" not in k]


def test_deep_supervision_layout(golden_keys):
    m = mau_b200.UrbanPredictor("unet++", 23, 828, 16, 8, 8, 32, 2, base_filters=8, deep_supervision=True)
    got = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert got == golden_keys["unetpp_ds_small"]


def test_ctor_errors_and_kwargs():
    with pytest.raises(ValueError):
        mau_b200.UrbanPredictor("resnet", *CTOR)
    # U-Net++ swallows the embedding flags like the reference (src/model.py:52-53)
    m = mau_b200.UrbanPredictor("unet++", *CTOR, temporal_embeddings=False, metadata_embeddings=False)
    assert m.model.conv0_1.conv1.weight.shape[1] == 64 + 128 + 128


def test_checkpoint_schema_roundtrip(tmp_path):
    """torch.save dict of src/train.py:305-316 -> load_state_dict(strict=True)."""
    mt, kw = VARIANTS["unet_metaemb"]
    torch.manual_seed(0)
    a = mau_b200.UrbanPredictor(mt, *CTOR, **kw)
    opt = torch.optim.AdamW(a.parameters(), lr=1e-4, weight_decay=1e-3)
    ck = {"epoch": 3, "step": 1234, "model_state_dict": a.state_dict(), "optimizer_state_dict": opt.state_dict(),
          "loss": 0.25, "hyperparameters": {"learning_rate": 1e-4, "batch_size": 16, "weight_decay": 1e-3,
                                            "temporal_dim": 64, "meta_dim": 64, "lstm_hidden": 96, "model_type": mt,
                                            "target_channels": 2, "input_channels": 23, **kw},
          "model_type": mt, "study_name": "s", "trial_id": 0, "metadata_input_length": 8}
    f = tmp_path / "ck.pth"
    torch.save(ck, f)
    ld = torch.load(f, weights_only=False)
    b = mau_b200.UrbanPredictor(ld["model_type"], 23, 828, ld["hyperparameters"]["temporal_dim"],
                                ld["metadata_input_length"], ld["hyperparameters"]["meta_dim"],
                                ld["hyperparameters"]["lstm_hidden"], 2,
                                temporal_embeddings=ld["hyperparameters"]["temporal_embeddings"],
                                metadata_embeddings=ld["hyperparameters"]["metadata_embeddings"])
    missing = b.load_state_dict(ld["model_state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k


def test_no_cpu_fallback():
    m = mau_b200.UrbanPredictor("unet", *CTOR)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 23, 32, 32), torch.zeros(1, 8), torch.zeros(1, 8))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mau_b200.h")).read()
    declared = set(re.findall(r"\b(mau_[a-z0-9_]+)\s*\(", hdr)) - {"mau_grad_ready_fn"}
    assert declared, "no declarations parsed"
    L = engine.lib()
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in include/mau_b200.h but not exported"
    assert declared == set(engine.EXPORTS)
    assert L.mau_version() >= 100


def _cfg(model_type, te, me, B=1, H=250, W=250, training=0, filters=(64, 128, 256, 512, 1024)):
    return dict(model_type=model_type, spatial_channels=23, temporal_dim=64, meta_features=8, meta_dim=64, lstm_dim=96,
                out_channels=2, filters=list(filters), temporal_embeddings=te, metadata_embeddings=me,
                deep_supervision=0, batch=B, height=H, width=W, seq_len=828, training=training, precision=0,
                device=0, flags=0)


@pytest.mark.parametrize("name,gflop", [("unet_noemb", 106.31), ("unet_metaemb", 106.57), ("unet_emb", 106.90),
                                        ("unetpp_emb", 319.03)])
def test_layer_graph_work_matches_survey(name, gflop, golden_keys):
    """Host logic without a GPU: state order, roles, per-tile algorithmic FLOPs."""
    mt, kw = VARIANTS[name]
    te = int(kw.get("temporal_embeddings", True)); me = int(kw.get("metadata_embeddings", True))
    d = engine.describe(_cfg(1 if mt == "unet++" else 0, te, me))
    assert abs(d["fwd_flops"] / 1e9 - gflop) < 0.01
    assert [s[0] for s in d["state"]] == [k for k, _, _ in golden_keys[name]]
    for (nm, numel, role), (_, shape, dtype) in zip(d["state"], golden_keys[name]):
        n = 1
        for s in shape:
            n *= s
        assert numel == n, nm
        assert (role == 3) == (dtype == "torch.int64")
        if "temporal_encoder" in nm:
            assert role == (0 if (te or mt == "unet++") else 1)
        if "meta_encoder" in nm:
            assert role == (0 if (me or mt == "unet++") else 1)


def test_layer_graph_shapes_unet():
    d = engine.describe(_cfg(0, 0, 1))
    L = {l["name"]: l for l in d["layers"]}
    assert (L["conv3_1.conv1"]["cin"], L["conv3_1.conv1"]["cout"], L["conv3_1.conv1"]["h"]) == (1536, 512, 31)
    assert (L["conv4_0.conv1"]["cin"], L["conv4_0.conv1"]["h"]) == (576, 15)
    assert abs(L["conv0_1.conv1"]["flops"] / 1e9 - 13.824) < 1e-3       # SURVEY.md per-layer table
    assert L["conv0_0.conv1"]["kp"] == 64                                # 23 channels padded to one K chunk
    # every torch.cat is a slice: the encoder output lands in the decoder's concat buffer
    assert L["conv0_0.conv2"]["out"] == "cat0" and L["conv0_0.conv2"]["out_c0"] == 0


def test_layer_graph_shapes_unetpp():
    d = engine.describe(_cfg(1, 1, 1))
    L = {l["name"]: l for l in d["layers"]}
    assert L["conv0_4.conv1"]["cin"] == 512 and L["conv0_4.conv1"]["segs"] == [[0, 256], [640, 128], [768, 128]]
    assert L["conv3_1.conv1"]["cin"] == 1664
    buf = {b["name"]: b for b in d["buffers"]}
    assert buf["level0"]["c"] == 4 * 64 + 4 * 128 + 128


def test_describe_rejects_bad_configs():
    with pytest.raises(RuntimeError):
        engine.describe(_cfg(7, 1, 1))
    with pytest.raises(RuntimeError):
        engine.describe(_cfg(0, 1, 1, H=8, W=8))


def test_plan_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = engine.make_config(_cfg(0, 0, 1))
    import ctypes as C
    h = C.c_void_p()
    rc = engine.lib().mau_plan_create(C.byref(c), C.byref(h))
    assert rc != 0 and b"no CUDA device" in engine.lib().mau_last_error()


def test_shared_maps_plan_runs_the_encoder_once():
    """MAU_FLAG_SHARED_MAPS (metadata-sensitivity sweep, reference test/metadata_sensitivity.py:294-311): every
    buffer up to the bottleneck input holds ONE tile, the decoder the whole batch; executed conv FLOPs drop by
    the encoder's share while the dense (reference) FLOPs stay the roofline numerator."""
    B = 50
    dense = engine.describe(_cfg(0, 1, 1, B=B))
    cfg = _cfg(0, 1, 1, B=B)
    cfg["flags"] = engine.FLAG_SHARED_MAPS
    sweep = engine.describe(cfg)
    assert sweep["fwd_flops"] == dense["fwd_flops"]
    L = {l["name"]: l for l in sweep["layers"]}
    for l in range(4):
        assert L[f"conv{l}_0.conv1"]["b"] == 1 and L[f"conv{l}_0.conv2"]["b"] == 1
    assert L["conv4_0.conv1"]["b"] == B and L["conv0_1.conv2"]["b"] == B
    buf = {b["name"]: b for b in sweep["buffers"]}
    assert buf["maps_nhwc"]["b"] == 1 and buf["x0_0.shared"]["b"] == 1 and buf["cat0"]["b"] == B
    enc = sum(L[f"conv{l}_0.conv{k}"]["flops"] for l in range(4) for k in (1, 2))          # dense FLOPs, B tiles
    assert abs(dense["exec_conv_flops"] - sweep["exec_conv_flops"] - enc * (B - 1) / B) < 1e-6 * dense["exec_conv_flops"]
    assert sweep["workspace_bytes"] < dense["workspace_bytes"]
    train = _cfg(0, 1, 1, B=4, training=1)
    train["flags"] = engine.FLAG_SHARED_MAPS
    with pytest.raises(RuntimeError, match="inference-only"):
        engine.describe(train)


def test_shared_maps_detection_is_host_logic():
    """Batch-expanded (stride-0) inputs select the sweep plan; materialised batches do not unless asserted."""
    m = mau_b200.UrbanPredictor("unet", *CTOR).eval()
    x1 = torch.zeros(1, 23, 32, 32)
    assert x1.expand(5, -1, -1, -1).stride(0) == 0 and x1.repeat(5, 1, 1, 1).stride(0) != 0
    assert m.model.shared_maps == "auto"
    m.assume_shared_maps(True)
    assert m.model.shared_maps is True
    with pytest.raises(ValueError):
        m.assume_shared_maps("sometimes")


def test_fused_adamw_has_no_cpu_fallback_and_torch_state_layout():
    p = torch.nn.Parameter(torch.zeros(4))
    opt = mau_b200.FusedAdamW([p], lr=1e-4, weight_decay=1e-3)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(4))], lr=1e-4, weight_decay=1e-3)
    assert opt.state_dict()["param_groups"][0].keys() == ref.state_dict()["param_groups"][0].keys()
    opt.step()                                    # no gradients: nothing to do, like the stock optimizer
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()


@pytest.mark.parametrize("name", list(VARIANTS))
def test_state_tensor_walk_matches_state_dict_order(name):
    """The engine receives the state as a pointer array in state_dict order; the module collects it with a direct
    walk over parameters / persistent buffers: it must be the state_dict's tensors, in its order, by identity."""
    mt, kw = VARIANTS[name]
    m = mau_b200.UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
    a = m.model._state_tensors()
    b = list(m.model.state_dict(keep_vars=True).values())
    assert len(a) == len(b) and all(x is y for x, y in zip(a, b))


def test_losses_mirror_has_the_reference_names_and_refuses_cpu_tensors():
    """mau_b200.losses mirrors src/utils/losses.py (same functions and arguments); like the model it has no CPU path."""
    import inspect
    from mau_b200 import losses
    want = {"gradient_loss": ["pred", "target"], "compute_loss_mse": ["outputs", "targets"],
            "compute_loss_mse_gradient": ["outputs", "targets", "lambda_grad"],
            "compute_loss_l1_grad_ssim": ["outputs", "targets", "lambda_grad", "lambda_ssim"],
            "compute_all_loss": ["outputs", "targets", "lambda_grad", "lambda_ssim"]}
    for name, args in want.items():
        assert list(inspect.signature(getattr(losses, name)).parameters) == args
    assert inspect.signature(losses.compute_loss_l1_grad_ssim).parameters["lambda_ssim"].default == 0.5
    x = torch.zeros(1, 2, 16, 16)
    for fn in (losses.compute_loss_mse, losses.compute_loss_mse_gradient, losses.compute_loss_l1_grad_ssim, losses.compute_all_loss,
               losses.gradient_loss):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn(x, x)


def test_metric_rows_reproduce_the_reference_row_list():
    """engine.metric_rows turns the device reductions into the rows of test/evaluate.py:239-275; here the reductions
    are formed with NumPy from the oracle's own class map so that the host logic is checked without a GPU."""
    import numpy as np
    from oracle import unet_oracle as O
    maps, _, _, tgt = O.synthetic_batch(2, 24, 20, T=4, seed=5)
    maps[1, 0:9] = 0
    maps[1, 6] = 1.0                                                     # sample 1: a single class present
    pred = torch.randn(2, 2, 24, 20, generator=torch.Generator().manual_seed(1))
    dw, ref_rows = O.eval_metrics(maps.numpy(), pred.numpy(), tgt.numpy(), temp_mean=14.5, temp_std=7.25)
    lap = O.laplacian_variance(pred.numpy(), tgt.numpy(), 14.5, 7.25)
    p, g = pred.numpy().copy(), tgt.numpy().copy()
    p[:, 1] = p[:, 1] * np.float32(7.25) + np.float32(14.5)
    g[:, 1] = g[:, 1] * np.float32(7.25) + np.float32(14.5)
    sums = np.zeros((2, 2, 10, 3))
    for i in range(2):
        for ch in range(2):
            d = (p[i, ch] - g[i, ch]).astype(np.float64)
            for slot in range(10):
                m = np.ones_like(dw[i], bool) if slot == 0 else dw[i] == slot - 1
                sums[i, ch, slot] = [m.sum(), np.abs(d[m]).sum(), (d[m] ** 2).sum()]
    rows = engine.metric_rows(torch.from_numpy(sums), torch.from_numpy(lap), first_sample_idx=40)
    assert len(rows) == len(ref_rows)
    for r, (i, ch, k, n, mae, rmse) in zip(rows, ref_rows):
        assert r["sample_idx"] == 40 + i and r["channel"] == ("after_ndvi", "after_temp")[ch]
        assert r["dw_class"] == ("overall" if k < 0 else engine.DW_CLASSES[k])
        assert abs(r["mae"] - mae) < 1e-5 * max(1, abs(mae)) and abs(r["rmse"] - rmse) < 1e-5 * max(1, abs(rmse))
        assert (r["laplacian_var_pred"] is not None) == (k < 0)
        if k < 0:
            assert r["laplacian_var_pred"] == lap[i, ch, 0] and r["laplacian_var_gt"] == lap[i, ch, 1]
    assert [r["dw_class"] for r in rows if r["sample_idx"] == 41 and r["channel"] == "after_ndvi"] == ["overall", "built"]
