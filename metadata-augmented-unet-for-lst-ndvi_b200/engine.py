"""ctypes binding of the C-ABI library (include/mau_b200.h) + the autograd bridge.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every
kernel that runs belongs to ``libmau_b200.so``.  If the library is missing the import of
this module still succeeds (so CPU-only tooling can introspect the package) but any attempt
to build a plan raises ``RuntimeError`` -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
import weakref
from typing import Dict, List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmau_b200.so")

MODEL_UNET, MODEL_UNETPP = 0, 1
PRECISIONS = {"bf16": 0, "fp32": 1}
FLAG_SHARED_MAPS, FLAG_CONV_TAPLOAD, FLAG_CONV_FFMA = 1, 2, 4
ROLE_PARAM, ROLE_UNUSED_PARAM, ROLE_RUNNING_STAT, ROLE_COUNTER = 0, 1, 2, 3


def default_precision() -> str:
    p = os.environ.get("MAU_PRECISION", "bf16")
    return p if p in PRECISIONS else "bf16"


class MauConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "model_type", "spatial_channels", "temporal_dim", "meta_features", "meta_dim", "lstm_dim",
        "out_channels")] + [("filters", C.c_int32 * 5)] + [(n, C.c_int32) for n in (
        "temporal_embeddings", "metadata_embeddings", "deep_supervision", "batch", "height", "width",
        "seq_len", "training", "precision", "device", "flags")]


GRAD_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int)
STATS_SYNC = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)

_lib = None
_lib_lock = threading.Lock()

# every symbol include/mau_b200.h declares (tests check the .so exports all of them)
EXPORTS = (
    "mau_last_error", "mau_version", "mau_launch_count", "mau_plan_create", "mau_plan_destroy",
    "mau_plan_workspace_bytes", "mau_plan_num_state", "mau_plan_state_info", "mau_plan_describe_config",
    "mau_plan_flops", "mau_plan_exec_flops", "mau_plan_forward", "mau_plan_forward_staged", "mau_plan_backward", "mau_plan_set_grad_hook", "mau_plan_wait_backward_streams", "mau_plan_set_stats_sync", "mau_plan_set_state_version",
    "mau_plan_buffer_ptr", "mau_plan_profile", "mau_plan_profile_read", "mau_loss_forward_backward", "mau_loss_backward", "mau_eval_metrics", "mau_laplacian_sums", "mau_ssim_work_floats", "mau_ssim_loss", "mau_ssim_forward", "mau_ssim_backward",
    "mau_op_conv3x3", "mau_op_conv3x3_dgrad", "mau_op_conv3x3_bench", "mau_op_conv3x3_wgrad", "mau_op_conv3x3_wgrad_bench", "mau_set_sm_reserve", "mau_op_bw_bench", "mau_adamw_step", "mau_cast_f32_bf16", "mau_cast_bf16_f32", "mau_op_maxpool2x2", "mau_op_bilinear", "mau_op_bilinear_bwd",
    "mau_op_nchw_to_nhwc", "mau_op_nhwc_to_nchw", "mau_op_lstm_last_hidden",
)


def lib():
    """Load libmau_b200.so once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"mau_b200: CUDA library not built ({LIB_PATH} missing). Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` -- there is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.mau_last_error.restype = C.c_char_p
        L.mau_version.restype = C.c_int
        L.mau_launch_count.restype = C.c_int64
        L.mau_plan_create.argtypes = [C.POINTER(MauConfig), C.POINTER(C.c_void_p)]
        L.mau_plan_destroy.argtypes = [C.c_void_p]
        L.mau_plan_workspace_bytes.argtypes = [C.c_void_p]
        L.mau_plan_workspace_bytes.restype = C.c_size_t
        L.mau_plan_num_state.argtypes = [C.c_void_p]
        L.mau_plan_state_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.mau_plan_describe_config.argtypes = [C.POINTER(MauConfig), C.c_char_p, C.c_size_t]
        L.mau_plan_flops.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mau_plan_exec_flops.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mau_plan_forward.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.mau_plan_forward_staged.argtypes = L.mau_plan_forward.argtypes
        L.mau_plan_backward.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
        L.mau_plan_set_grad_hook.argtypes = [C.c_void_p, GRAD_HOOK, C.c_void_p]
        L.mau_plan_set_state_version.argtypes = [C.c_void_p, C.c_uint64]
        L.mau_plan_wait_backward_streams.argtypes = [C.c_void_p, C.c_void_p]
        L.mau_plan_buffer_ptr.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.mau_plan_set_stats_sync.argtypes = [C.c_void_p, STATS_SYNC, C.c_void_p, C.c_int]
        L.mau_plan_profile.argtypes = [C.c_void_p, C.c_int]
        L.mau_plan_profile_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_float),
                                            C.c_int, C.POINTER(C.c_int)]
        L.mau_loss_forward_backward.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mau_loss_backward.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mau_eval_metrics.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.mau_laplacian_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                         C.c_void_p, C.c_void_p]
        L.mau_ssim_work_floats.argtypes = [C.c_int, C.c_int, C.c_int]
        L.mau_ssim_work_floats.restype = C.c_int64
        L.mau_ssim_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.mau_ssim_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.mau_ssim_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        L.mau_op_conv3x3.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p]
        L.mau_op_conv3x3_dgrad.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mau_op_conv3x3_bench.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mau_op_conv3x3_wgrad.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p]
        L.mau_op_conv3x3_wgrad_bench.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_void_p]
        L.mau_set_sm_reserve.argtypes = [C.c_int]
        L.mau_adamw_step.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_int64, C.c_void_p]
        L.mau_cast_f32_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.mau_cast_bf16_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.mau_op_bw_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]
        L.mau_op_maxpool2x2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
        L.mau_op_bilinear.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p]
        L.mau_op_bilinear_bwd.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mau_op_nchw_to_nhwc.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p]
        L.mau_op_nhwc_to_nchw.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p]
        L.mau_op_lstm_last_hidden.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


# ---------------------------------------------------------------------- #
# State epoch.  Eval plans keep packed bf16 weights / folded BatchNorm vectors while the caller's state is unchanged.
# torch's per-tensor ``_version`` counters see every in-place torch op, but NOT writes through raw device pointers:
# FusedAdamW.step (csrc/optim.cu) and the training forward's running_mean / running_var / num_batches_tracked updates.
# Both bump this process-wide epoch, which is folded into the version handed to ``mau_plan_set_state_version``.
# ---------------------------------------------------------------------- #
_state_epoch = 0


def bump_state_epoch() -> int:
    """Call after writing parameters or buffers through raw device pointers (outside torch's version tracking)."""
    global _state_epoch
    _state_epoch += 1
    return _state_epoch


def state_epoch() -> int:
    return _state_epoch


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mau_last_error()
        raise RuntimeError(f"mau_b200 {what}: {msg.decode() if msg else 'error'} (code {rc})")


def _stream_ptr() -> C.c_void_p:
    # the *calling thread's* current stream (backward runs on the autograd worker thread)
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_config(cfg: Dict) -> MauConfig:
    c = MauConfig()
    for k, v in cfg.items():
        if k == "filters":
            c.filters = (C.c_int32 * 5)(*[int(x) for x in v])
        else:
            setattr(c, k, int(v))
    return c


def describe(cfg: Dict) -> Dict:
    """Layer graph of a configuration as a dict (host logic only; needs no GPU)."""
    c = make_config(cfg)
    buf = C.create_string_buffer(1 << 20)
    check(lib().mau_plan_describe_config(C.byref(c), buf, len(buf)), "describe")
    return json.loads(buf.value.decode())


def is_staged_maps(maps: torch.Tensor) -> bool:
    """bf16 tensors are STAGED tiles: NHWC, channel stride a multiple of 8 (see :func:`stage_maps`)."""
    return maps.dtype == torch.bfloat16


def stage_maps(maps: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 NCHW tiles [B, C, H, W] (the reference's contract, src/dataset.py:99-106) -> the engine's bf16 NHWC input layout
    [B, H, W, round_up(C, 8)] (pad channels zero), on whatever device ``maps`` lives on.  A producer that stages on the
    HOST (into a pinned ``out``) halves the PCIe bytes per tile; ``model(staged, series, metadata)`` then skips the layout
    kernel.  Round-to-nearest-even like the kernel, so outputs are bit-identical to passing the fp32 tiles."""
    B, Cc, H, W = maps.shape
    cs = (Cc + 7) // 8 * 8
    if out is None:
        out = torch.zeros((B, H, W, cs), dtype=torch.bfloat16, device=maps.device)
    elif out.shape != (B, H, W, cs) or out.dtype != torch.bfloat16:
        raise ValueError(f"out must be bf16 [{B},{H},{W},{cs}]")
    out[..., :Cc].copy_(maps.permute(0, 2, 3, 1))
    return out


class _DevMem:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_alias(ptr: int, n: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    if dtype == torch.bfloat16:       # no bf16 in the array interface: alias as 16-bit words, reinterpret
        return torch.as_tensor(_DevMem(ptr, n, "<i2"), device=device).view(torch.bfloat16)
    typestr = {torch.float64: "<f8", torch.float32: "<f4"}[dtype]
    return torch.as_tensor(_DevMem(ptr, n, typestr), device=device)


class Plan:
    """One (variant, B, H, W, T, mode, precision) instance of the engine."""

    def __init__(self, cfg: Dict):
        self.cfg = dict(cfg)
        self._c = make_config(cfg)
        self._h = C.c_void_p()
        self.device = torch.device("cuda", int(cfg.get("device", 0)))
        with torch.cuda.device(self.device):
            check(lib().mau_plan_create(C.byref(self._c), C.byref(self._h)), "plan_create")
        L = lib()
        self.num_state = L.mau_plan_num_state(self._h)
        self.roles: List[int] = []
        self.numels: List[int] = []
        for i in range(self.num_state):
            n, r = C.c_int64(), C.c_int()
            check(L.mau_plan_state_info(self._h, i, C.byref(n), C.byref(r)), "state_info")
            self.roles.append(r.value)
            self.numels.append(n.value)
        self.uses_metadata = bool(cfg.get("metadata_embeddings", 1))
        self.uses_series = bool(cfg.get("temporal_embeddings", 1))
        self._hook_ref = None
        self._pending = None        # token of the forward whose backward has not run yet (see HotPathFn)
        B, H, W = cfg["batch"], cfg["height"], cfg["width"]
        self.out_shape = ((4,) if cfg.get("deep_supervision") else ()) + (B, cfg["out_channels"], H, W)

    # ------------------------------------------------------------------ #
    def close(self):
        if self._h:
            lib().mau_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def pending(self) -> bool:
        """True while an autograd graph still needs this plan's saved activations."""
        return self._pending is not None and self._pending() is not None

    @property
    def workspace_bytes(self) -> int:
        return int(lib().mau_plan_workspace_bytes(self._h))

    def flops(self):
        f, b = C.c_double(), C.c_double()
        check(lib().mau_plan_flops(self._h, C.byref(f), C.byref(b)), "flops")
        return f.value, b.value

    def exec_conv_flops(self) -> float:
        """FLOPs the conv kernels execute per forward (dense conv FLOPs unless this is a shared-maps plan)."""
        return self.exec_flops()[0]

    def exec_flops(self):
        """(forward, backward) FLOPs the conv / dgrad / wgrad kernels of this plan execute per step."""
        f, b = C.c_double(), C.c_double()
        check(lib().mau_plan_exec_flops(self._h, C.byref(f), C.byref(b)), "exec_flops")
        return f.value, b.value

    def used_state_indices(self) -> List[int]:
        return [i for i, r in enumerate(self.roles) if r == ROLE_PARAM]

    def _ptr_array(self, tensors: Sequence[Optional[torch.Tensor]]):
        arr = (C.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = None if t is None else t.data_ptr()
        return arr

    def _check_state(self, state: Sequence[torch.Tensor]):
        if len(state) != self.num_state:
            raise RuntimeError(f"mau_b200: plan expects {self.num_state} state tensors, got {len(state)}")
        for i, t in enumerate(state):
            if t.numel() != self.numels[i] or t.device != self.device or not t.is_contiguous():
                raise RuntimeError(f"mau_b200: state tensor {i} mismatch: numel {t.numel()} vs "
                                   f"{self.numels[i]}, device {t.device} vs {self.device}")
            want = torch.int64 if self.roles[i] == ROLE_COUNTER else torch.float32
            if t.dtype != want:
                raise RuntimeError(f"mau_b200: state tensor {i} must be {want}, got {t.dtype} "
                                   "(keep the module in fp32; precision is chosen by set_precision)")

    def forward(self, state, maps, series, md, out: Optional[torch.Tensor] = None, staged: bool = False) -> torch.Tensor:
        key = tuple(t.data_ptr() for t in state)
        if key != getattr(self, "_validated", None):       # same memory as last time: shapes / dtypes were checked then
            self._check_state(state)
            self._validated = key
        if not self.cfg.get("training"):
            # in-place modification counters + the epoch of raw-pointer writes: unchanged state => packed weights are reused
            lib().mau_plan_set_state_version(self._h, 1 + sum(t._version for t in state) + (_state_epoch << 32))
        else:
            bump_state_epoch()          # this forward updates the BatchNorm running statistics through raw pointers
        if out is None:
            out = torch.empty(self.out_shape, device=self.device, dtype=torch.float32)
        fn = lib().mau_plan_forward_staged if staged else lib().mau_plan_forward
        with torch.cuda.device(self.device):
            check(fn(self._h, self._ptr_array(state), maps.data_ptr(), series.data_ptr() if series.numel() else None,
                     md.data_ptr() if md.numel() else None, out.data_ptr(), _stream_ptr()), "forward")
        return out

    def backward(self, grad_out: torch.Tensor, grads: Sequence[Optional[torch.Tensor]]):
        with torch.cuda.device(self.device):
            check(lib().mau_plan_backward(self._h, grad_out.data_ptr(), self._ptr_array(grads),
                                          _stream_ptr()), "backward")

    def buffer(self, name: str, grad: bool = False) -> Optional[torch.Tensor]:
        """Zero-copy [B, H, W, channel_stride] view of one of the plan's NHWC buffers (tests / tooling): the activation
        or, with ``grad=True``, its gradient twin (``None`` before the first backward)."""
        info = self._buffer_info().get(name)
        if info is None:
            raise KeyError(name)
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        check(lib().mau_plan_buffer_ptr(self._h, name.encode(), int(grad), C.byref(ptr), C.byref(nbytes)), "buffer_ptr")
        if not ptr.value:
            return None
        dtype = torch.float32 if self.cfg.get("precision") == PRECISIONS["fp32"] else torch.bfloat16
        n = info["b"] * info["h"] * info["w"] * info["cs"]
        return device_alias(ptr.value, n, dtype, self.device).view(info["b"], info["h"], info["w"], info["cs"])

    def _buffer_info(self):
        if getattr(self, "_binfo", None) is None:
            self._binfo = {b["name"]: b for b in self.describe()["buffers"]}
        return self._binfo

    def describe(self) -> Dict:
        if getattr(self, "_desc", None) is None:
            self._desc = describe(self.cfg)
        return self._desc

    def set_grad_hook(self, fn):
        """fn(first_index, last_index) is called from inside backward when those state
        gradients are final on the stream (data-parallel bucket launch point)."""
        if fn is None:
            self._hook_ref = None
            check(lib().mau_plan_set_grad_hook(self._h, C.cast(None, GRAD_HOOK), None), "set_grad_hook")
            return
        self._hook_ref = GRAD_HOOK(lambda _u, a, b: fn(a, b))
        check(lib().mau_plan_set_grad_hook(self._h, self._hook_ref, None), "set_grad_hook")

    def wait_backward_streams(self, stream_ptr: int):
        """Make the CUDA stream ``stream_ptr`` wait for the weight-gradient launches enqueued so far on the plan's own
        second stream (data-parallel: the communication stream calls this before reducing a bucket)."""
        check(lib().mau_plan_wait_backward_streams(self._h, C.c_void_p(stream_ptr)), "wait_backward_streams")

    def set_stats_sync(self, fn, world_size: int = 1):
        """SyncBN: fn(tensor) must all-reduce (SUM) the float64 device tensor in place on the current stream;
        it is called from inside forward / backward once per BatchNorm layer."""
        if fn is None:
            self._sync_ref = None
            check(lib().mau_plan_set_stats_sync(self._h, C.cast(None, STATS_SYNC), None, 1), "set_stats_sync")
            return
        dev = self.device

        def cb(_user, ptr, n):
            fn(device_alias(ptr, n, torch.float64, dev))
        self._sync_ref = STATS_SYNC(cb)
        check(lib().mau_plan_set_stats_sync(self._h, self._sync_ref, None, int(world_size)), "set_stats_sync")

    def profile(self, enable=True):
        check(lib().mau_plan_profile(self._h, int(enable)), "profile")

    def profile_read(self):
        names = C.create_string_buffer(1 << 16)
        ms = (C.c_float * 1024)()
        n = C.c_int()
        check(lib().mau_plan_profile_read(self._h, names, len(names), ms, 1024, C.byref(n)), "profile_read")
        nm = names.value.decode().split("\n")[: n.value]
        return list(zip(nm, [ms[i] for i in range(n.value)]))


class _PendingToken:
    """Marks a plan as holding the saved activations of a live autograd graph.  Released by backward, or when the graph
    is dropped without a backward (the ctx, and with it this token, is garbage-collected)."""

    __slots__ = ("plan", "__weakref__")

    def __init__(self, plan: "Plan"):
        self.plan = plan
        plan._pending = weakref.ref(self)      # weak: the graph (ctx) is the only owner, so dropping the graph frees the plan

    def release(self):
        if self.plan is not None and self.plan._pending is not None and self.plan._pending() is self:
            self.plan._pending = None
        self.plan = None

    def __del__(self):  # pragma: no cover - exercised through gc
        try:
            self.release()
        except Exception:
            pass


class HotPathFn(torch.autograd.Function):
    """The single autograd node of the model (reference: the whole nn.Module graph of
    src/model.py:261-292 / :123-193 as recorded by PyTorch autograd)."""

    @staticmethod
    def forward(ctx, plan: Plan, state, diff_idx, dp, maps, series, md, *diff_params):
        if plan.pending:
            raise RuntimeError("mau_b200: this plan still holds the activations of a forward that awaits its backward")
        out = plan.forward(state, maps, series, md, staged=is_staged_maps(maps))
        ctx.token = _PendingToken(plan)
        ctx.plan, ctx.diff_idx, ctx.n_state, ctx.dp = plan, diff_idx, len(state), dp
        # backward re-reads the series (LSTM BPTT) and the metadata (MLP) through the raw pointers the
        # plan kept from this forward: keep the tensors alive until then
        ctx.keep_alive = (maps, series, md, state)
        ctx.shapes = [p.shape for p in diff_params]
        ctx.params = diff_params if dp is not None else None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        plan: Plan = ctx.plan
        if plan._pending is None or plan._pending() is not ctx.token:
            raise RuntimeError("mau_b200: backward() called twice on the same graph, or the plan's saved activations were "
                               "released (a plan keeps the activations of exactly one forward)")
        grad_out = grad_out.contiguous().float()
        if ctx.dp is not None:        # data parallel: grads are views of one flat buffer, all-reduced
            grads_full, outs = ctx.dp.make_grads(plan, ctx.diff_idx, ctx.shapes, ctx.params)   # from inside backward
            plan.backward(grad_out, grads_full)
            ctx.dp.finish(plan)
            ctx.token.release()
            return (None, None, None, None, None, None, None, *outs)
        grads_full: List[Optional[torch.Tensor]] = [None] * ctx.n_state
        outs = []
        for i, shp in zip(ctx.diff_idx, ctx.shapes):
            g = torch.empty(shp, device=plan.device, dtype=torch.float32)
            grads_full[i] = g
            outs.append(g)
        plan.backward(grad_out, grads_full)
        ctx.token.release()
        return (None, None, None, None, None, None, None, *outs)


# ---------------------------------------------------------------------- #
# stand-alone operators of the path (loss terms, evaluation metrics)
# ---------------------------------------------------------------------- #
def _dev_f32(t: torch.Tensor, name: str, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The kernels read raw device pointers: insist on a CUDA tensor (on the same device as ``like``) and hand back a
    contiguous fp32 view of it (a no-op for the tensors the reference's callers hold)."""
    if not t.is_cuda:
        raise RuntimeError(f"mau_b200: {name} must be a CUDA tensor (no CPU fallback), got device {t.device}")
    if like is not None and t.device != like.device:
        raise RuntimeError(f"Expected all tensors to be on the same device: {name} is on {t.device}, expected {like.device}")
    return t.contiguous().float()


def loss_terms(pred: torch.Tensor, target: torch.Tensor, kind: str = "l1", lambda_grad: float = 0.1,
               need_grad: bool = True):
    """L1|MSE + lambda * gradient-difference loss and its gradient w.r.t. pred in one pass
    (reference src/utils/losses.py:5-25,33,67-70).  Returns (losses[4] tensor, grad or None)."""
    pred, target = _dev_f32(pred, "pred"), _dev_f32(target, "target", pred)
    if pred.dim() != 4 or target.shape != pred.shape:
        raise RuntimeError(f"pred and target must both be [B,C,H,W], got {tuple(pred.shape)} and {tuple(target.shape)}")
    B, Cc, H, W = pred.shape
    losses = torch.empty(4, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if need_grad else None
    with torch.cuda.device(pred.device):
        check(lib().mau_loss_forward_backward({"l1": 0, "mse": 1}[kind], pred.data_ptr(), target.data_ptr(),
                                              B, Cc, H, W, lambda_grad, losses.data_ptr(),
                                              grad.data_ptr() if need_grad else None, _stream_ptr()), "loss")
    return losses, grad


def loss_backward_terms(pred, target, kind, lambda_total, g_total=None, g_pixel=None, g_grad=None) -> torch.Tensor:
    """d(g_total*total + g_pixel*pixel + g_grad*gradient)/d pred (total = pixel + lambda_total*gradient); the g_* are 0-d
    device tensors or None.  One launch of ``loss_kernel`` (csrc/loss.cu) with the scalars read on the device."""
    pred, target = _dev_f32(pred, "pred"), _dev_f32(target, "target", pred)
    B, Cc, H, W = pred.shape
    grad = torch.empty_like(pred)
    gs = [None if g is None else _dev_f32(g, "upstream gradient", pred) for g in (g_total, g_pixel, g_grad)]
    with torch.cuda.device(pred.device):
        check(lib().mau_loss_backward({"l1": 0, "mse": 1}[kind], pred.data_ptr(), target.data_ptr(), B, Cc, H, W,
                                      float(lambda_total), *[None if g is None else g.data_ptr() for g in gs],
                                      grad.data_ptr(), _stream_ptr()), "loss_backward")
    return grad


class _LossFn(torch.autograd.Function):
    """(total, pixel, gradient) of src/utils/losses.py, each a differentiable 0-d tensor.  Forward: one pass for the three
    sums.  Backward: one pass that folds autograd's upstream scalars (device-side, no host sync, no separate scaling
    kernel) into d/d pred; entries nobody back-propagates through cost nothing (their upstream gradient arrives as None)."""

    @staticmethod
    def forward(ctx, pred, target, kind, lambda_grad):
        pred_c, target_c = pred.contiguous(), target.contiguous()
        losses, _ = loss_terms(pred_c, target_c, kind, lambda_grad, need_grad=False)
        ctx.set_materialize_grads(False)
        ctx.kind, ctx.lambda_grad = kind, float(lambda_grad)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(pred_c, target_c)
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_total, g_pixel, g_grad):
        pred, target = ctx.saved_tensors
        return loss_backward_terms(pred, target, ctx.kind, ctx.lambda_grad, g_total, g_pixel, g_grad), None, None, None


def compute_loss_l1_grad(outputs, targets, lambda_grad=0.1):
    """{'total','pixel','gradient'} like the L1+gradient part of the reference's
    compute_loss_l1_grad_ssim (src/utils/losses.py:59-99); every entry is differentiable."""
    total, pixel, grad = _LossFn.apply(outputs, targets, "l1", lambda_grad)
    return {"total": total, "pixel": pixel, "gradient": grad}


def compute_loss_mse_gradient(outputs, targets, lambda_grad=0.1):
    """src/utils/losses.py:41-57."""
    total, mse, grad = _LossFn.apply(outputs, targets, "mse", lambda_grad)
    return {"total": total, "mse": mse, "gradient": grad}


def eval_metrics(maps: torch.Tensor, pred: torch.Tensor, target: torch.Tensor,
                 temp_mean: float = 0.0, temp_std: float = 0.0):
    """Device version of test/evaluate.py:210-275.  Returns (dw_map int64 [B,H,W],
    sums float64 [B,C,10,3] = {count, sum|d|, sum d^2} for overall + 9 DW classes)."""
    pred = _dev_f32(pred, "pred")
    target, maps = _dev_f32(target, "target", pred), _dev_f32(maps, "maps", pred)
    if pred.dim() != 4 or target.shape != pred.shape or maps.dim() != 4 or maps.shape[0] != pred.shape[0] \
            or maps.shape[2:] != pred.shape[2:] or maps.shape[1] < 9:
        raise RuntimeError(f"eval_metrics expects maps [B,>=9,H,W] and pred / target [B,C,H,W], got {tuple(maps.shape)}, "
                           f"{tuple(pred.shape)}, {tuple(target.shape)}")
    B, Cc, H, W = pred.shape
    dw = torch.empty(B, H, W, device=pred.device, dtype=torch.int64)
    sums = torch.zeros(B, Cc, 10, 3, device=pred.device, dtype=torch.float64)
    with torch.cuda.device(pred.device):
        check(lib().mau_eval_metrics(maps.data_ptr(), maps.shape[1], pred.data_ptr(), target.data_ptr(),
                                     B, Cc, H, W, temp_mean, temp_std, dw.data_ptr(), sums.data_ptr(),
                                     _stream_ptr()), "eval_metrics")
    return dw, sums


def laplacian_variance(pred: torch.Tensor, target: torch.Tensor, temp_mean: float = 0.0, temp_std: float = 0.0) -> torch.Tensor:
    """``np.var(scipy.ndimage.laplace(x))`` of the un-normalised prediction and target of every (sample, channel)
    (test/evaluate.py:241-242: ``laplacian_var_pred`` / ``laplacian_var_gt``).  Returns float64 [B, C, 2]."""
    pred = _dev_f32(pred, "pred")
    target = _dev_f32(target, "target", pred)
    if pred.dim() != 4 or target.shape != pred.shape:
        raise RuntimeError(f"pred and target must both be [B,C,H,W], got {tuple(pred.shape)} and {tuple(target.shape)}")
    B, Cc, H, W = pred.shape
    sums = torch.empty(B, Cc, 4, device=pred.device, dtype=torch.float64)
    with torch.cuda.device(pred.device):
        check(lib().mau_laplacian_sums(pred.data_ptr(), target.data_ptr(), B, Cc, H, W, temp_mean, temp_std,
                                       sums.data_ptr(), _stream_ptr()), "laplacian_sums")
    n = float(H * W)
    mean = sums[..., 0::2] / n
    return sums[..., 1::2] / n - mean * mean


def ssim_loss_terms(pred: torch.Tensor, target: torch.Tensor, need_grad: bool = True):
    """``1 - piq.ssim(scaled pred, scaled target, data_range=1, reduction='none').mean()`` of
    src/utils/losses.py:72-90 and its gradient w.r.t. ``pred``.  Returns (loss[1] tensor, grad or None)."""
    pred = _dev_f32(pred, "pred")
    target = _dev_f32(target, "target", pred)
    if pred.dim() != 4 or target.shape != pred.shape:
        raise RuntimeError(f"pred and target must both be [B,C,H,W], got {tuple(pred.shape)} and {tuple(target.shape)}")
    B, Cc, H, W = pred.shape
    n_work = int(lib().mau_ssim_work_floats(B, H, W))
    work = torch.empty(max(n_work, 1), device=pred.device, dtype=torch.float32)
    acc = torch.empty(1, device=pred.device, dtype=torch.float64)
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if need_grad else None
    with torch.cuda.device(pred.device):
        check(lib().mau_ssim_loss(pred.data_ptr(), target.data_ptr(), B, Cc, H, W, loss.data_ptr(),
                                  grad.data_ptr() if need_grad else None, work.data_ptr(), acc.data_ptr(), _stream_ptr()),
              "ssim_loss")
    return loss, grad


def ssim_forward_terms(pred: torch.Tensor, target: torch.Tensor):
    """Forward half: (loss[1], work) -- ``work`` holds the per-window derivative maps the backward half needs."""
    pred = _dev_f32(pred, "pred")
    target = _dev_f32(target, "target", pred)
    if pred.dim() != 4 or target.shape != pred.shape:
        raise RuntimeError(f"pred and target must both be [B,C,H,W], got {tuple(pred.shape)} and {tuple(target.shape)}")
    B, Cc, H, W = pred.shape
    work = torch.empty(max(int(lib().mau_ssim_work_floats(B, H, W)), 1), device=pred.device, dtype=torch.float32)
    acc = torch.empty(1, device=pred.device, dtype=torch.float64)
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    with torch.cuda.device(pred.device):
        check(lib().mau_ssim_forward(pred.data_ptr(), target.data_ptr(), B, Cc, H, W, loss.data_ptr(), work.data_ptr(),
                                     acc.data_ptr(), _stream_ptr()), "ssim_forward")
    return loss, work


def ssim_backward_terms(pred: torch.Tensor, target: torch.Tensor, work: torch.Tensor, upstream: Optional[torch.Tensor]):
    """Backward half: upstream * d loss / d pred (``upstream``: 0-d device tensor or None), one kernel launch."""
    pred = _dev_f32(pred, "pred")
    target = _dev_f32(target, "target", pred)
    B, Cc, H, W = pred.shape
    grad = torch.empty_like(pred)
    up = None if upstream is None else _dev_f32(upstream, "upstream gradient", pred)
    with torch.cuda.device(pred.device):
        check(lib().mau_ssim_backward(pred.data_ptr(), target.data_ptr(), B, Cc, H, W, work.data_ptr(),
                                      None if up is None else up.data_ptr(), grad.data_ptr(), _stream_ptr()), "ssim_backward")
    return grad


class _SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred_c, target_c = pred.contiguous(), target.contiguous()
        loss, work = ssim_forward_terms(pred_c, target_c)
        if ctx.needs_input_grad[0]:               # validation runs under no_grad (src/train.py:33): nothing is kept
            ctx.save_for_backward(pred_c, target_c)
            ctx.work = work
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        return ssim_backward_terms(pred, target, ctx.work, g), None


def ssim_loss(outputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """Differentiable scalar ``ssim_loss`` of src/utils/losses.py:88-89 (parity with piq unpinned)."""
    return _SsimFn.apply(outputs, targets)


# Dynamic World class names, reference src/utils/visualization.py:5-8
DW_CLASSES = {0: "water", 1: "trees", 2: "grass", 3: "flooded_vegetation", 4: "crops", 5: "shrub_and_scrub", 6: "built",
              7: "bare", 8: "snow_and_ice"}


def metric_rows(sums, lap_var=None, channel_names=("after_ndvi", "after_temp"), first_sample_idx: int = 0):
    """The per-sample rows test/evaluate.py:239-275 appends, from the device reductions: ``sums`` [B,C,10,3] of
    :func:`eval_metrics` and (optionally) ``lap_var`` [B,C,2] of :func:`laplacian_variance` (tensors on any device or
    arrays).  One ``overall`` row per (sample, channel) followed by one row per Dynamic World class present in the
    sample, in the reference's order; keys ``sample_idx, channel, dw_class, mae, rmse, laplacian_var_pred,
    laplacian_var_gt`` (the caller adds its city / date columns).  One device-to-host copy of a few kilobytes replaces
    the reference's copy of the full prediction, target and input tensors (test/evaluate.py:188)."""
    import numpy as np
    s = sums.detach().cpu().numpy() if isinstance(sums, torch.Tensor) else np.asarray(sums)
    lv = None if lap_var is None else (lap_var.detach().cpu().numpy() if isinstance(lap_var, torch.Tensor) else np.asarray(lap_var))
    B, Cc = s.shape[:2]
    if len(channel_names) < Cc:
        raise ValueError(f"{Cc} channels but only {len(channel_names)} channel names")
    rows = []
    for i in range(B):
        for ch in range(Cc):
            for slot in range(10):
                n = s[i, ch, slot, 0]
                if n <= 0:
                    continue            # the reference skips classes absent from the sample (np.any(mask))
                overall = slot == 0
                rows.append({"sample_idx": first_sample_idx + i, "channel": channel_names[ch],
                             "dw_class": "overall" if overall else DW_CLASSES[slot - 1],
                             "mae": float(s[i, ch, slot, 1] / n), "rmse": float(np.sqrt(s[i, ch, slot, 2] / n)),
                             "laplacian_var_pred": float(lv[i, ch, 0]) if overall and lv is not None else None,
                             "laplacian_var_gt": float(lv[i, ch, 1]) if overall and lv is not None else None})
    return rows
