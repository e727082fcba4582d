"""CPU oracle for the Metadata-Augmented U-Net hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg use it, and there only as the checker / CPU baseline.

It is a *functional* restatement (plain ``torch.nn.functional`` on CPU, fp32) of
the algorithm in the reference's ``src/model.py`` driven directly from a
``state_dict`` -- it does not import or copy the reference.  Each function cites
the reference lines it follows (paths relative to the reference repo root).

Parity status: pinned.  ``oracle/gen_golden.py`` imports the real reference in
the build container and writes ``tests/golden/*.npz``; ``tests/test_oracle.py``
checks this restatement against those vectors.  The SSIM loss term
(``piq.ssim``, third-party, unpinned, absent) is *not* restated: parity unpinned
for that term only.

``emulate_bf16=True`` (``forward`` / ``train_step_grads``) is the SAME graph with
values rounded to bfloat16 exactly where the B200 engine stores bf16:
the NHWC input, the packed convolution weights, every stored pre-BatchNorm output
``z``, every stored activation ``y`` (BN+ReLU, bilinear stages, embedding planes),
and -- in backward -- every stored gradient (``dz``, activation gradients).  All
arithmetic between two stores stays fp32, like the kernels (fp32 accumulation in
TMEM, fp32 BatchNorm coefficients, fp32 LSTM / MLP / head / loss).  With the flag off
nothing changes, so the pin above also pins the graph of the emulation; the rounding
points themselves are a statement about the engine, not about the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5        # nn.BatchNorm2d default, src/model.py:13,15
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default


# --------------------------------------------------------------------------- #
# bf16 storage emulation (off by default)
# --------------------------------------------------------------------------- #
_EMULATE_BF16 = False


class emulate_bf16_storage:
    """Context manager: inside it the building blocks below round to bfloat16 where the engine stores bf16."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        global _EMULATE_BF16
        self.prev, _EMULATE_BF16 = _EMULATE_BF16, self.on
        return self

    def __exit__(self, *exc):
        global _EMULATE_BF16
        _EMULATE_BF16 = self.prev
        return False


def _bf16(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(torch.float32)      # round-to-nearest-even, like __float2bfloat16_rn


class _StoreBf16(torch.autograd.Function):
    """A tensor written to HBM as bf16: the value is rounded in forward, its gradient (also stored as bf16 by the
    engine) is rounded in backward."""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _PackBf16(torch.autograd.Function):
    """Packed convolution weights: bf16 in forward; the weight gradient leaves the wgrad kernel in fp32."""

    @staticmethod
    def forward(ctx, w):
        return _bf16(w)

    @staticmethod
    def backward(ctx, g):
        return g


def _st(x: Tensor) -> Tensor:
    return _StoreBf16.apply(x) if _EMULATE_BF16 else x


def _pk(w: Tensor) -> Tensor:
    return _PackBf16.apply(w) if _EMULATE_BF16 else w


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def _bn(sd: Dict[str, Tensor], prefix: str, x: Tensor, training: bool,
        new_stats: Optional[Dict[str, Tensor]]) -> Tensor:
    """nn.BatchNorm2d (src/model.py:13,15): batch statistics (biased variance)
    when training, running statistics otherwise; running update uses the
    unbiased variance with momentum 0.1."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not training:
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)
    rm2, rv2 = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm2, rv2, w, b, True, BN_MOMENTUM, BN_EPS)
    if new_stats is not None:
        new_stats[prefix + ".running_mean"] = rm2
        new_stats[prefix + ".running_var"] = rv2
        new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    return y


def vgg_block(sd, prefix: str, x: Tensor, training: bool, new_stats=None) -> Tensor:
    """VGGBlock.forward, src/model.py:18-21: relu(bn1(conv1(x))), relu(bn2(conv2(.)))."""
    # bf16 emulation: training plans store the pre-BN output z and the activation y; eval plans fold BatchNorm into
    # the convolution epilogue (fp32 accumulator -> scale/shift -> ReLU -> one bf16 store)
    zst = _st if training else (lambda t: t)
    x = zst(F.conv2d(x, _pk(sd[prefix + ".conv1.weight"]), sd[prefix + ".conv1.bias"], padding=1))
    x = _st(F.relu(_bn(sd, prefix + ".bn1", x, training, new_stats)))
    x = zst(F.conv2d(x, _pk(sd[prefix + ".conv2.weight"]), sd[prefix + ".conv2.bias"], padding=1))
    x = _st(F.relu(_bn(sd, prefix + ".bn2", x, training, new_stats)))
    return x


def lstm_last_hidden(sd, prefix: str, series: Tensor) -> Tensor:
    """nn.LSTM(1, hidden, batch_first=True) over the whole (zero padded) series,
    zero initial state, last hidden state (src/model.py:26,29-33).
    Gate order i, f, g, o (PyTorch convention)."""
    w_ih, w_hh = sd[prefix + ".weight_ih_l0"], sd[prefix + ".weight_hh_l0"]
    b_ih, b_hh = sd[prefix + ".bias_ih_l0"], sd[prefix + ".bias_hh_l0"]
    B, T = series.shape
    Hd = w_hh.shape[1]
    h = series.new_zeros(B, Hd)
    c = series.new_zeros(B, Hd)
    for t in range(T):
        x_t = series[:, t:t + 1]                                  # (B,1)
        gates = x_t @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


def temporal_encoder(sd, prefix: str, series: Tensor) -> Tensor:
    """TemporalEncoder.forward, src/model.py:29-34."""
    h = lstm_last_hidden(sd, prefix + ".lstm", series)
    return F.linear(h, sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"])


def metadata_encoder(sd, prefix: str, md: Tensor) -> Tensor:
    """MetadataEncoder.forward, src/model.py:41-48: Linear-ReLU-Linear."""
    x = F.relu(F.linear(md, sd[prefix + ".fc.0.weight"], sd[prefix + ".fc.0.bias"]))
    return F.linear(x, sd[prefix + ".fc.2.weight"], sd[prefix + ".fc.2.bias"])


def bilinear_ac(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """Bilinear resize, align_corners=True (src/model.py:121, 219, 245)."""
    return _st(F.interpolate(x, size=size, mode="bilinear", align_corners=True))


def head(sd, prefix: str, x: Tensor) -> Tensor:
    """1x1 conv + tanh on channel 0 when there are two channels, src/model.py:284-292."""
    out = F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"])
    if out.shape[1] == 2:
        return torch.cat([torch.tanh(out[:, 0:1]), out[:, 1:2]], dim=1)
    return out


# --------------------------------------------------------------------------- #
# the two networks
# --------------------------------------------------------------------------- #
def unet_forward(sd, maps, series, md, *, temporal_embeddings=True, metadata_embeddings=True,
                 training=False, new_stats=None, p="model") -> Tensor:
    """UrbanPredictor_unet.forward, src/model.py:261-292."""
    t_emb = temporal_encoder(sd, p + ".temporal_encoder", series) if temporal_embeddings else None
    m_emb = metadata_encoder(sd, p + ".meta_encoder", md) if metadata_embeddings else None
    pool = lambda t: F.max_pool2d(t, 2, 2)
    blk = lambda name, t: vgg_block(sd, f"{p}.{name}", t, training, new_stats)

    x0_0 = blk("conv0_0", _st(maps))
    x1_0 = blk("conv1_0", pool(x0_0))
    x2_0 = blk("conv2_0", pool(x1_0))
    x3_0 = blk("conv3_0", pool(x2_0))
    x4_0 = pool(x3_0)
    # fuse_embeddings, src/model.py:248-259: temporal first, then metadata
    B, _, H, W = x4_0.shape
    cat = [x4_0]
    if t_emb is not None:
        cat.append(_st(t_emb[:, :, None, None].expand(B, t_emb.shape[1], H, W)))
    if m_emb is not None:
        cat.append(_st(m_emb[:, :, None, None].expand(B, m_emb.shape[1], H, W)))
    x4_0 = blk("conv4_0", torch.cat(cat, 1) if len(cat) > 1 else x4_0)

    def up_to(src, ref):
        # self.up (x2, src/model.py:219) then _upsample_match (src/model.py:243-246)
        u = bilinear_ac(src, (src.shape[2] * 2, src.shape[3] * 2))
        if u.shape[2:] != ref.shape[2:]:
            u = bilinear_ac(u, tuple(ref.shape[2:]))
        return u

    x3_1 = blk("conv3_1", torch.cat([x3_0, up_to(x4_0, x3_0)], 1))
    x2_1 = blk("conv2_1", torch.cat([x2_0, up_to(x3_1, x2_0)], 1))
    x1_1 = blk("conv1_1", torch.cat([x1_0, up_to(x2_1, x1_0)], 1))
    x0_1 = blk("conv0_1", torch.cat([x0_0, up_to(x1_1, x0_0)], 1))
    return head(sd, p + ".final", x0_1)


def unetpp_forward(sd, maps, series, md, *, deep_supervision=False, training=False,
                   new_stats=None, p="model"):
    """UrbanPredictor_unetpp.forward, src/model.py:123-193 (always both embeddings)."""
    t_emb = temporal_encoder(sd, p + ".temporal_encoder", series)
    m_emb = metadata_encoder(sd, p + ".meta_encoder", md)
    emb = _st(torch.cat([t_emb, m_emb], 1))                       # src/model.py:103 (bf16 planes in the level buffers)
    pool = lambda t: F.max_pool2d(t, 2, 2)
    blk = lambda name, t: vgg_block(sd, f"{p}.{name}", t, training, new_stats)

    def node(name, same_level: Sequence[Tensor], lower: Tensor):
        H, W = same_level[0].shape[2:]
        e = emb[:, :, None, None].expand(emb.shape[0], emb.shape[1], H, W)
        return blk(name, torch.cat([*same_level, bilinear_ac(lower, (H, W)), e], 1))

    x0_0 = blk("conv0_0", _st(maps))
    x1_0 = blk("conv1_0", pool(x0_0))
    x0_1 = node("conv0_1", [x0_0], x1_0)
    x2_0 = blk("conv2_0", pool(x1_0))
    x1_1 = node("conv1_1", [x1_0], x2_0)
    x0_2 = node("conv0_2", [x0_0, x0_1], x1_1)
    x3_0 = blk("conv3_0", pool(x2_0))
    x2_1 = node("conv2_1", [x2_0], x3_0)
    x1_2 = node("conv1_2", [x1_0, x1_1], x2_1)
    x0_3 = node("conv0_3", [x0_0, x0_1, x0_2], x1_2)
    x4_0 = blk("conv4_0", pool(x3_0))
    x3_1 = node("conv3_1", [x3_0], x4_0)
    x2_2 = node("conv2_2", [x2_0, x2_1], x3_1)
    x1_3 = node("conv1_3", [x1_0, x1_1, x1_2], x2_2)
    x0_4 = node("conv0_4", [x0_0, x0_1, x0_2, x0_3], x1_3)
    if deep_supervision:                                          # src/model.py:180-185, no tanh
        return [F.conv2d(x, sd[f"{p}.final{i}.weight"], sd[f"{p}.final{i}.bias"])
                for i, x in zip((1, 2, 3, 4), (x0_1, x0_2, x0_3, x0_4))]
    return head(sd, p + ".final", x0_4)


def forward(sd, model_type: str, maps, series, md, emulate_bf16: bool = False, **kw):
    """UrbanPredictor.forward dispatch, src/model.py:295-329.  ``emulate_bf16``: see the module docstring."""
    if emulate_bf16:
        with emulate_bf16_storage(True):
            return forward(sd, model_type, maps, series, md, **kw)
    if model_type == "unet":
        kw.pop("deep_supervision", None)
        return unet_forward(sd, maps, series, md, **kw)
    if model_type == "unet++":
        kw.pop("temporal_embeddings", None)
        kw.pop("metadata_embeddings", None)   # swallowed by **kwargs, src/model.py:52-53
        return unetpp_forward(sd, maps, series, md, **kw)
    raise ValueError(f"Unsupported model_type: {model_type}")


# --------------------------------------------------------------------------- #
# losses (src/utils/losses.py) -- SSIM term intentionally not restated
# --------------------------------------------------------------------------- #
def gradient_loss(pred: Tensor, target: Tensor) -> Tensor:
    """src/utils/losses.py:5-25."""
    dy_p = (pred[:, :, 1:, :] - pred[:, :, :-1, :]).abs()
    dx_p = (pred[:, :, :, 1:] - pred[:, :, :, :-1]).abs()
    dy_t = (target[:, :, 1:, :] - target[:, :, :-1, :]).abs()
    dx_t = (target[:, :, :, 1:] - target[:, :, :, :-1]).abs()
    return (dy_p - dy_t).abs().mean() + (dx_p - dx_t).abs().mean()


def loss_mse_gradient(out: Tensor, tgt: Tensor, lambda_grad=0.1) -> Dict[str, Tensor]:
    """src/utils/losses.py:41-57."""
    mse = F.mse_loss(out, tgt)
    g = gradient_loss(out, tgt)
    return {"total": mse + lambda_grad * g, "mse": mse, "gradient": g}


def loss_l1_gradient(out: Tensor, tgt: Tensor, lambda_grad=0.1) -> Dict[str, Tensor]:
    """The L1 + gradient part of compute_loss_l1_grad_ssim, src/utils/losses.py:67-70,95
    (the piq SSIM term of :88-92 is left to piq; parity unpinned for it)."""
    l1 = F.l1_loss(out, tgt)
    g = gradient_loss(out, tgt)
    return {"total": l1 + lambda_grad * g, "pixel": l1, "gradient": g}


# --------------------------------------------------------------------------- #
# evaluation metrics (test/evaluate.py:210-275), NumPy like the reference
# --------------------------------------------------------------------------- #
def dw_class_map(input_stack_i):
    """test/evaluate.py:212-217: argmax_c(input[c]*c) over the 9 Dynamic World channels.
    int64 index map; ties -> lowest index (np.argmax)."""
    import numpy as np
    return np.argmax(np.stack([input_stack_i[c] * c for c in range(9)]), axis=0)


def eval_metrics(input_stack, pred, tgt, temp_mean=None, temp_std=None):
    """Per sample / channel: overall MAE, RMSE (test/evaluate.py:239-240) and the
    masked MAE / RMSE per DW class when the class is present (:259-263).  Channel 1
    (temperature) is un-normalised first when statistics are given (:23-41).
    Returns (dw_maps[int64 B,H,W], rows) with rows = list of
    (sample, channel, cls or -1, count, mae, rmse)."""
    import numpy as np
    pred = np.array(pred, dtype=np.float32, copy=True)
    tgt = np.array(tgt, dtype=np.float32, copy=True)
    if temp_mean is not None:
        pred[:, 1] = pred[:, 1] * np.float32(temp_std) + np.float32(temp_mean)
        tgt[:, 1] = tgt[:, 1] * np.float32(temp_std) + np.float32(temp_mean)
    rows, maps = [], []
    for i in range(pred.shape[0]):
        dw = dw_class_map(np.asarray(input_stack[i]))
        maps.append(dw)
        for ch in range(pred.shape[1]):
            p, g = pred[i, ch], tgt[i, ch]
            rows.append((i, ch, -1, p.size, float(np.mean(np.abs(p - g))),
                         float(np.sqrt(np.mean((p - g) ** 2)))))
            for k in range(9):
                m = dw == k
                if np.any(m):
                    rows.append((i, ch, k, int(m.sum()), float(np.mean(np.abs(p[m] - g[m]))),
                                 float(np.sqrt(np.mean((p[m] - g[m]) ** 2)))))
    return np.stack(maps), rows


def laplacian_variance(pred, tgt, temp_mean=None, temp_std=None):
    """``np.var(laplace(x))`` of every un-normalised (sample, channel) plane of pred and tgt
    (test/evaluate.py:241-242; ``scipy.ndimage.laplace``, default mode 'reflect' = edge value
    repeated).  Restated without scipy: per axis the second difference [1,-2,1] in double rounded
    to fp32, the two axes added in fp32.  Returns float64 [B, C, 2] = (var_pred, var_gt).
    Pinned against scipy itself in tests/test_oracle.py."""
    import numpy as np
    pred = np.array(pred, dtype=np.float32, copy=True)
    tgt = np.array(tgt, dtype=np.float32, copy=True)
    if temp_mean is not None and pred.shape[1] > 1:      # channel 1 is the temperature (conf/config.yaml:29)
        pred[:, 1] = pred[:, 1] * np.float32(temp_std) + np.float32(temp_mean)
        tgt[:, 1] = tgt[:, 1] * np.float32(temp_std) + np.float32(temp_mean)

    def lap(a):
        p = np.pad(a.astype(np.float64), 1, mode="edge")
        d2y = (p[:-2, 1:-1] + p[2:, 1:-1] - 2.0 * p[1:-1, 1:-1]).astype(np.float32)
        d2x = (p[1:-1, :-2] + p[1:-1, 2:] - 2.0 * p[1:-1, 1:-1]).astype(np.float32)
        return d2y + d2x

    out = np.zeros(pred.shape[:2] + (2,), np.float64)
    for i in range(pred.shape[0]):
        for ch in range(pred.shape[1]):
            out[i, ch, 0] = np.var(lap(pred[i, ch]).astype(np.float64))
            out[i, ch, 1] = np.var(lap(tgt[i, ch]).astype(np.float64))
    return out


# --------------------------------------------------------------------------- #
# training step (src/train.py:244-256) -- gradients through torch autograd
# --------------------------------------------------------------------------- #
def train_step_grads(sd, model_type, maps, series, md, tgt, loss="l1", lambda_grad=0.1, emulate_bf16=False, **kw):
    """fwd (train-mode BN) + loss + bwd.  Returns (out, loss, grads{name}, new_stats).  ``emulate_bf16``: the same step
    with bf16 rounding at the engine's storage points (forward values and backward gradients), see the module docstring."""
    if emulate_bf16:
        with emulate_bf16_storage(True):      # autograd runs the custom backward functions inside lv.backward() below
            return train_step_grads(sd, model_type, maps, series, md, tgt, loss=loss, lambda_grad=lambda_grad, **kw)
    params = {k: v.detach().clone().requires_grad_(True)
              for k, v in sd.items() if v.is_floating_point() and "running_" not in k}
    full = dict(sd)
    full.update(params)
    new_stats: Dict[str, Tensor] = {}
    out = forward(full, model_type, maps, series, md, training=True, new_stats=new_stats, **kw)
    if loss == "l1":
        lv = F.l1_loss(out, tgt)
    elif loss == "mse":
        lv = F.mse_loss(out, tgt)
    elif loss == "abs_mean":
        lv = out.abs().mean()
    elif loss == "l1_grad":
        lv = loss_l1_gradient(out, tgt, lambda_grad)["total"]
    elif loss == "mse_grad":
        lv = loss_mse_gradient(out, tgt, lambda_grad)["total"]
    else:
        raise ValueError(loss)
    lv.backward()
    grads = {k: v.grad for k, v in params.items()}
    return out.detach(), lv.detach(), grads, new_stats


# --------------------------------------------------------------------------- #
# synthetic inputs
# --------------------------------------------------------------------------- #
def synthetic_batch(B: int, H: int, W: int, T: int = 828, seed: int = 1002, shared_maps=False):
    """Seeded synthetic tiles: one-hot DW channels 0-8 and 14-22, N(0,1) RGB/LST,
    U(-1,1) NDVI; N(0,1) series; 4 x N(0,1) metadata ++ raw [2019,7,2023,7]."""
    g = torch.Generator().manual_seed(seed)
    nb = 1 if shared_maps else B
    cls = torch.randint(0, 9, (nb, H, W), generator=g)
    maps = torch.zeros(nb, 23, H, W)
    maps[:, 0:9] = F.one_hot(cls, 9).permute(0, 3, 1, 2).float()
    maps[:, 9:12] = torch.randn(nb, 3, H, W, generator=g)
    maps[:, 12] = torch.rand(nb, H, W, generator=g) * 2 - 1
    maps[:, 13] = torch.randn(nb, H, W, generator=g)
    redraw = torch.rand(nb, H, W, generator=g) < 0.1
    cls2 = torch.where(redraw, torch.randint(0, 9, (nb, H, W), generator=g), cls)
    maps[:, 14:23] = F.one_hot(cls2, 9).permute(0, 3, 1, 2).float()
    series = torch.randn(nb, T, generator=g)
    md = torch.cat([torch.randn(B, 4, generator=g),
                    torch.tensor([2019., 7., 2023., 7.]).expand(B, 4)], 1)
    tgt = torch.stack([torch.rand(B, H, W, generator=g) * 2 - 1,
                       torch.randn(B, H, W, generator=g)], 1)
    if shared_maps:
        maps = maps.expand(B, -1, -1, -1).contiguous()
        series = series.expand(B, -1).contiguous()
    return maps, series, md, tgt


def perturb_bn_stats(sd, seed: int = 5):
    """Eval configs: running_mean ~ N(0,.1), running_var ~ U(.5,1.5) so that BN
    folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        if k.endswith("running_mean"):
            sd[k].copy_(torch.randn(sd[k].shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            sd[k].copy_(torch.rand(sd[k].shape, generator=g) + 0.5)
    return sd
