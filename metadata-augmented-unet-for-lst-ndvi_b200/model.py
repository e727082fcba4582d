"""Drop-in ``UrbanPredictor`` whose forward/backward run on the B200 engine.

Boundary mirrored: reference ``src/model.py:295-329`` (``UrbanPredictor``), with the
constructor signature, sub-module names, parameter registration order and buffers of
``UrbanPredictor_unet`` (``src/model.py:195-241``) and ``UrbanPredictor_unetpp``
(``src/model.py:51-96``), so ``state_dict()`` is key-for-key / shape-for-shape equal and
``optimizer.state_dict()`` parameter indices line up.

The ``torch.nn`` layers below are *parameter containers only* (they give the reference's
default initialisation and key names); none of their ``forward`` methods is ever called.
All arithmetic happens in the C-ABI library (``csrc/``) through ``engine.Plan``.
There is no CPU / eager fallback: tensors must live on a CUDA device and the library must
be built, otherwise a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from . import engine

try:  # the reference logs its embedding flags at construction (src/model.py:202)
    from loguru import logger as _log
except Exception:  # pragma: no cover - loguru is optional here
    _log = None


class _ConvPair(nn.Module):
    """Parameter holder for one reference ``VGGBlock`` (src/model.py:9-16):
    conv1/bn1/conv2/bn2, 3x3 convolutions with bias, BatchNorm2d with running stats."""

    def __init__(self, cin: int, cmid: int, cout: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cmid, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(cmid)
        self.conv2 = nn.Conv2d(cmid, cout, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(cout)


class _SeriesParams(nn.Module):
    """Holder for ``TemporalEncoder`` (src/model.py:23-27): lstm(1->hidden) + fc."""

    def __init__(self, hidden: int, out_dim: int):
        super().__init__()
        self.lstm = nn.LSTM(input_size=1, hidden_size=hidden, batch_first=True)
        self.fc = nn.Linear(hidden, out_dim)


class _MetaParams(nn.Module):
    """Holder for ``MetadataEncoder`` (src/model.py:38-45): fc.0 (in->32), fc.2 (32->out)."""

    def __init__(self, in_features: int, out_dim: int):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(in_features, 32), nn.ReLU(), nn.Linear(32, out_dim))


class _EngineNet(nn.Module):
    """Common machinery: plan cache + dispatch into the C-ABI engine."""

    _variant = -1

    def _engine_config(self) -> Dict[str, int]:
        raise NotImplementedError

    def _init_engine_state(self):
        # not registered as parameters/buffers -> invisible to state_dict()
        object.__setattr__(self, "_plans", {})
        object.__setattr__(self, "_dp", None)
        object.__setattr__(self, "precision", engine.default_precision())
        object.__setattr__(self, "shared_maps", "auto")

    # ------------------------------------------------------------------ #
    def _state_tensors(self) -> List[torch.Tensor]:
        """The tensors of ``state_dict()`` in its order (per module: parameters, persistent buffers, then the
        sub-modules in registration order) without building the prefixed key dictionary (4x cheaper per call)."""
        out: List[torch.Tensor] = []

        def walk(mod):
            for p in mod._parameters.values():
                if p is not None:
                    out.append(p)
            for k, b in mod._buffers.items():
                if b is not None and k not in mod._non_persistent_buffers_set:
                    out.append(b)
            for c in mod._modules.values():
                if c is not None:
                    walk(c)
        walk(self)
        return out

    def _plan_for(self, B: int, H: int, W: int, T: int, training: bool, device: torch.device, shared: bool = False):
        """Plan cache.  A plan holds ONE set of saved activations, so a training plan whose forward still waits for its
        backward (``l1 = crit(model(a)); l2 = crit(model(b)); (l1 + l2).backward()``) is never reused: the second forward
        of that shape takes the next free slot (at most ``_MAX_PENDING`` graphs of one shape alive at a time)."""
        base = (B, H, W, T, bool(training), self.precision, device.index, bool(shared))
        for slot in range(self._MAX_PENDING):
            plan = self._plans.get(base + (slot,))
            if plan is None or not plan.pending:
                break
        else:
            raise RuntimeError(f"mau_b200: {self._MAX_PENDING} forward passes of shape {(B, H, W)} are waiting for their "
                               "backward; each holds a full set of saved activations -- call backward() or drop the outputs")
        key = base + (slot,)
        if plan is None:
            cfg = dict(self._engine_config())
            cfg.update(batch=B, height=H, width=W, seq_len=T, training=int(training),
                       precision=engine.PRECISIONS[self.precision],
                       device=device.index if device.index is not None else torch.cuda.current_device())
            cfg["flags"] = int(cfg.get("flags", 0)) | int(os.environ.get("MAU_FLAGS", "0"))   # debug knobs
            if shared:
                cfg["flags"] |= engine.FLAG_SHARED_MAPS
            plan = engine.Plan(cfg)
            if len(self._plans) >= 8:           # bound the workspace held by stale shapes (never a plan autograd still needs)
                for k in list(self._plans):
                    if not self._plans[k].pending:
                        self._plans.pop(k).close()
                        break
            self._plans[key] = plan
        return plan

    _MAX_PENDING = 4

    def assume_shared_maps(self, mode="auto"):
        """``True``: the caller guarantees that all rows of ``maps`` / ``temp_series`` are identical (the
        metadata-sensitivity sweep) -- the encoder and the LSTM then run once per forward.  ``"auto"``
        (default): only batch-expanded (stride-0) inputs are treated that way.  ``False``: never."""
        if mode not in (True, False, "auto"):
            raise ValueError("mode must be True, False or 'auto'")
        object.__setattr__(self, "shared_maps", mode)
        return self

    def release_plans(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()

    def forward(self, maps, temp_series, metadata):
        if not maps.is_cuda:
            raise RuntimeError("mau_b200: the hot path runs on a CUDA device only (no CPU fallback); "
                               "move the model and its inputs to cuda")
        staged = engine.is_staged_maps(maps)       # bf16 NHWC tiles produced by engine.stage_maps (halves the H2D bytes)
        Cs = self._cfg["spatial_channels"]
        if staged:
            if maps.dim() != 4 or maps.shape[3] != (Cs + 7) // 8 * 8:
                raise RuntimeError(f"staged maps must be bf16 [B,H,W,{(Cs + 7) // 8 * 8}], got {tuple(maps.shape)}")
            if self.precision != "bf16":
                raise RuntimeError("staged bf16 maps need the bf16 engine (set_precision('bf16'))")
            B, H, W, _ = maps.shape
        else:
            if maps.dim() != 4 or maps.shape[1] != Cs:
                raise RuntimeError(f"maps must be [B,{Cs},H,W], got {tuple(maps.shape)}")
            B, _, H, W = maps.shape
        if min(H, W) < 16:
            raise RuntimeError("tile edge must be >= 16 (four 2x2 poolings)")
        T = int(temp_series.shape[1]) if temp_series.dim() == 2 else 0
        if not self._cfg["temporal_embeddings"]:
            T = 0            # the series is ignored (src/model.py:263): do not key plans (and their workspaces) on its length
        # Sensitivity sweep (reference test/metadata_sensitivity.py:294-311): every row carries the same tile
        # and series, only the metadata differs.  Detected for free when the caller passes batch-expanded
        # views (stride 0 on dim 0, e.g. ``x.expand(50, -1, -1, -1)``), or asserted by the caller with
        # ``assume_shared_maps(True)`` for materialised ``.repeat()`` batches.  Eval mode only.
        shared = False
        if not self.training and B > 1 and self.shared_maps is not False:
            expanded = maps.stride(0) == 0 and (temp_series.dim() != 2 or temp_series.stride(0) == 0
                                                or not self._cfg["temporal_embeddings"])
            shared = bool(self.shared_maps is True or expanded)
        if shared:
            maps, temp_series = maps[:1], temp_series[:1]
        plan = self._plan_for(B, H, W, T, self.training, maps.device, shared)
        if self.training and getattr(self, "_dp", None) is not None:
            self._dp.prepare_plan(plan)          # SyncBN hook (data-parallel training)
        # the engine reads raw device pointers: an input left on the host (or on another GPU) must fail here, like
        # PyTorch's "Expected all tensors to be on the same device", not as an illegal address inside a kernel
        for name, t, used in (("temp_series", temp_series, plan.uses_series), ("metadata", metadata, plan.uses_metadata)):
            if used and t.device != maps.device:
                raise RuntimeError(f"Expected all tensors to be on the same device: maps is on {maps.device}, {name} on {t.device}")
        if plan.uses_series and (temp_series.dim() != 2 or T < 1 or temp_series.shape[0] != (1 if shared else B)):
            raise RuntimeError(f"temp_series must be [{B},T] with T >= 1, got {tuple(temp_series.shape)}")
        maps = maps.contiguous() if staged else maps.contiguous().float()
        # an ignored argument (flag off, src/model.py:263-264) never crosses the ABI, wherever it lives
        temp_series = temp_series.contiguous().float() if plan.uses_series else maps.new_empty(0)
        metadata = metadata.contiguous().float() if plan.uses_metadata else maps.new_empty(0)
        if metadata.dim() == 2 and metadata.shape[0] != B and plan.uses_metadata:
            raise RuntimeError(f"metadata must have {B} rows, got {metadata.shape[0]}")
        uses_meta = plan.uses_metadata
        if uses_meta and (metadata.dim() != 2 or metadata.shape[1] != self._cfg["meta_features"]):
            raise RuntimeError(f"metadata must be [B,{self._cfg['meta_features']}], got {tuple(metadata.shape)}")
        state = self._state_tensors()
        need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in state)
        if not need_grad:
            return self._finish(plan.forward(state, maps, temp_series, metadata, staged=staged))
        if not self.training:
            raise RuntimeError("mau_b200: backward through an eval()-mode forward is not supported; "
                               "call model.train() or wrap inference in torch.no_grad()")
        used = plan.used_state_indices()
        diff = [i for i in used if state[i].requires_grad]
        out = engine.HotPathFn.apply(plan, state, diff, getattr(self, "_dp", None), maps, temp_series,
                                     metadata, *[state[i] for i in diff])
        return self._finish(out)

    def _finish(self, out):
        if self._cfg.get("deep_supervision"):
            return [out[i] for i in range(out.shape[0])]
        return out


class UrbanPredictor_unet(_EngineNet):
    """State layout of reference ``UrbanPredictor_unet`` (src/model.py:195-241)."""

    def __init__(self, spatial_channels, seq_len, temporal_dim, meta_features, meta_dim, lstm_dim,
                 out_channels, nb_filter=None, temporal_embeddings=True, metadata_embeddings=True):
        super().__init__()
        if _log is not None:
            _log.info(f"UrbanPredictor_unet[b200] temporal_embeddings={temporal_embeddings}, "
                      f"metadata_embeddings={metadata_embeddings}")
        f = list(nb_filter) if nb_filter is not None else [32, 64, 128, 256, 512]
        self.temporal_dim, self.meta_dim = temporal_dim, meta_dim
        self.temporal_embeddings, self.metadata_embeddings = temporal_embeddings, metadata_embeddings
        # registration order is the contract (optimizer state indexes parameters by it)
        self.temporal_encoder = _SeriesParams(lstm_dim, temporal_dim)
        self.meta_encoder = _MetaParams(meta_features, meta_dim)
        self.conv0_0 = _ConvPair(spatial_channels, f[0], f[0])
        self.conv1_0 = _ConvPair(f[0], f[1], f[1])
        self.conv2_0 = _ConvPair(f[1], f[2], f[2])
        self.conv3_0 = _ConvPair(f[2], f[3], f[3])
        fused = f[3] + (temporal_dim if temporal_embeddings else 0) + (meta_dim if metadata_embeddings else 0)
        self.conv4_0 = _ConvPair(fused, f[4], f[4])
        self.conv3_1 = _ConvPair(f[3] + f[4], f[3], f[3])
        self.conv2_1 = _ConvPair(f[2] + f[3], f[2], f[2])
        self.conv1_1 = _ConvPair(f[1] + f[2], f[1], f[1])
        self.conv0_1 = _ConvPair(f[0] + f[1], f[0], f[0])
        self.final = nn.Conv2d(f[0], out_channels, kernel_size=1)
        object.__setattr__(self, "_cfg", dict(
            model_type=engine.MODEL_UNET, spatial_channels=spatial_channels, temporal_dim=temporal_dim,
            meta_features=meta_features, meta_dim=meta_dim, lstm_dim=lstm_dim, out_channels=out_channels,
            filters=f, temporal_embeddings=int(bool(temporal_embeddings)),
            metadata_embeddings=int(bool(metadata_embeddings)), deep_supervision=0))
        self._init_engine_state()

    def _engine_config(self):
        return self._cfg


class UrbanPredictor_unetpp(_EngineNet):
    """State layout of reference ``UrbanPredictor_unetpp`` (src/model.py:51-96).  Unknown
    keyword arguments are swallowed like the reference does (src/model.py:52-53), so both
    embeddings are always on."""

    def __init__(self, spatial_channels, seq_len, temporal_dim, meta_features, meta_dim, lstm_dim,
                 out_channels, base_filters=32, deep_supervision=False, **kwargs):
        super().__init__()
        f = [base_filters * m for m in (1, 2, 4, 8, 16)]
        e = temporal_dim + meta_dim
        self.deep_supervision = deep_supervision
        self.embed_dim = e
        for lvl in range(5):                                   # encoders conv0_0 .. conv4_0
            cin = spatial_channels if lvl == 0 else f[lvl - 1]
            setattr(self, f"conv{lvl}_0", _ConvPair(cin, f[lvl], f[lvl]))
        for depth in (1, 2, 3, 4):                             # conv0_1..conv3_1, conv0_2.., conv0_4
            for lvl in range(0, 5 - depth):
                setattr(self, f"conv{lvl}_{depth}", _ConvPair(f[lvl] * depth + f[lvl + 1] + e, f[lvl], f[lvl]))
        self.temporal_encoder = _SeriesParams(lstm_dim, temporal_dim)
        self.meta_encoder = _MetaParams(meta_features, meta_dim)
        if deep_supervision:
            for i in (1, 2, 3, 4):
                setattr(self, f"final{i}", nn.Conv2d(f[0], out_channels, kernel_size=1))
        else:
            self.final = nn.Conv2d(f[0], out_channels, kernel_size=1)
        object.__setattr__(self, "_cfg", dict(
            model_type=engine.MODEL_UNETPP, spatial_channels=spatial_channels, temporal_dim=temporal_dim,
            meta_features=meta_features, meta_dim=meta_dim, lstm_dim=lstm_dim, out_channels=out_channels,
            filters=f, temporal_embeddings=1, metadata_embeddings=1,
            deep_supervision=int(bool(deep_supervision))))
        self._init_engine_state()

    def _engine_config(self):
        return self._cfg


class UrbanPredictor(nn.Module):
    """Same constructor / forward / state_dict surface as reference ``UrbanPredictor``
    (src/model.py:295-329).  ``model_type`` in {'unet', 'unet++'}; anything else raises
    ``ValueError`` (src/model.py:326)."""

    def __init__(self, model_type, spatial_channels, seq_len, temporal_dim, meta_features, meta_dim,
                 lstm_dim, out_channels, base_filters=64, deep_supervision=False, **kwargs):
        super().__init__()
        common = dict(spatial_channels=spatial_channels, seq_len=seq_len, temporal_dim=temporal_dim,
                      meta_features=meta_features, meta_dim=meta_dim, lstm_dim=lstm_dim,
                      out_channels=out_channels)
        if model_type == "unet++":
            self.model = UrbanPredictor_unetpp(base_filters=base_filters,
                                               deep_supervision=deep_supervision, **common, **kwargs)
        elif model_type == "unet":
            self.model = UrbanPredictor_unet(nb_filter=[base_filters * m for m in (1, 2, 4, 8, 16)],
                                             **common, **kwargs)
        else:
            raise ValueError(f"Unsupported model_type: {model_type}")

    def forward(self, maps, temp_series, metadata):
        return self.model(maps, temp_series, metadata)

    # convenience (not part of the reference surface)
    def forward_sweep(self, maps, temp_series, metadata):
        """One tile, many metadata rows (reference test/metadata_sensitivity.py:294-311 builds this with
        ``.repeat(50, ...)``): ``maps`` [1,C,H,W], ``temp_series`` [1,T], ``metadata`` [B,F] -> [B,out,H,W].
        Identical to ``forward(maps.repeat(B,1,1,1), temp_series.repeat(B,1), metadata)`` in eval mode."""
        B = metadata.shape[0]
        return self.model(maps[:1].expand(B, -1, -1, -1), temp_series[:1].expand(B, -1), metadata)

    def assume_shared_maps(self, mode="auto"):
        self.model.assume_shared_maps(mode)
        return self

    def set_precision(self, precision: str):
        """'bf16' (tcgen05 tensor-core path, default) or 'fp32' (FFMA path, 1e-5 parity mode)."""
        if precision not in engine.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(engine.PRECISIONS)}")
        self.model.precision = precision
        return self
