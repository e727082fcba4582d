"""Data-parallel training across the GPUs of one NVLink/NVSwitch box: one process per GPU
(``torchrun``), the batch is sharded across ranks, and the parameter-gradient all-reduce is
launched from *inside* backward, layer by layer, as each layer's weight gradient becomes final
(``mau_plan_set_grad_hook``), on a side stream so that it overlaps the remaining backward kernels.

The reference is single-GPU (``CONFIG.device = "cuda:0"``, src/train.py:99); this is a new
capability.  BatchNorm uses per-rank batch statistics by default (the throughput mode; running statistics
can be averaged across ranks with :meth:`DataParallel.sync_buffers` before a checkpoint).  With
``sync_bn=True`` every BatchNorm all-reduces its per-channel sums (2*C doubles, forward and backward)
so that an N-rank step equals the single-process step at the global batch -- the reference semantics.

Inference needs none of this: tiles are independent in eval mode, shard them and run replicas.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import engine


class GradReducer:
    """Averages slices of one flat gradient buffer across ranks as they become ready.

    Device-agnostic core (also exercised with gloo on CPU in tests/test_parallel.py):
    ``ready(a, b)`` marks elements [a, b) final; ranges are merged while contiguous and flushed
    once ``bucket_numel`` elements are pending; ``finish()`` flushes the rest and waits."""

    def __init__(self, flat: torch.Tensor, group=None, bucket_numel: int = 1 << 20, grad_dtype: str = "fp32"):
        if grad_dtype not in ("fp32", "bf16"):
            raise ValueError("grad_dtype must be 'fp32' or 'bf16'")
        self.flat, self.group, self.bucket_numel = flat, group, bucket_numel
        # bf16 wire format: the bucket is rounded to bf16, averaged in bf16 by the collective and widened back; halves
        # the NVLink payload and the time the collective's CTAs compete with the backward kernels (fp32 = exact average)
        self.wire = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device) if grad_dtype == "bf16" else None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending: List[Tuple[int, int]] = []
        self.cuda = flat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat.device) if self.cuda else None
        self.launched = 0
        self.use_avg = self.cuda          # NCCL has ReduceOp.AVG; gloo does not
        self.extra_wait = None            # callable(stream_ptr): producers that run on streams torch does not know

    def _launch(self, a: int, b: int):
        if self.world == 1 or b <= a:
            return
        sl = self.flat[a:b]
        self.launched += 1
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            self.comm_stream.wait_event(ev)
            if self.extra_wait is not None:       # the plan's weight-gradient stream (engine.Plan.wait_backward_streams)
                self.extra_wait(self.comm_stream.cuda_stream)
            with torch.cuda.stream(self.comm_stream):
                if self.wire is None:
                    dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    w = self.wire[a:b]
                    st = torch.cuda.current_stream(self.flat.device).cuda_stream
                    with torch.cuda.device(self.flat.device):
                        engine.check(engine.lib().mau_cast_f32_bf16(sl.data_ptr(), w.data_ptr(), b - a, 16, st), "cast")
                        dist.all_reduce(w, op=dist.ReduceOp.AVG, group=self.group)
                        engine.check(engine.lib().mau_cast_bf16_f32(w.data_ptr(), sl.data_ptr(), b - a, 16, st), "cast")
        elif self.wire is None:
            dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.group)
            sl.div_(self.world)
        else:                    # host / gloo path of the unit tests: same rounding points with torch ops
            w = self.wire[a:b]
            w.copy_(sl)
            w32 = w.float()
            dist.all_reduce(w32, op=dist.ReduceOp.SUM, group=self.group)
            sl.copy_((w32 / self.world).to(torch.bfloat16))

    def ready(self, a: int, b: int):
        if self.pending and self.pending[-1][0] == b:          # extends the last range downwards
            self.pending[-1] = (a, self.pending[-1][1])
        elif self.pending and self.pending[-1][1] == a:        # ... or upwards
            self.pending[-1] = (self.pending[-1][0], b)
        else:
            self.pending.append((a, b))
        if sum(y - x for x, y in self.pending) >= self.bucket_numel:
            self.flush()

    def flush(self):
        for a, b in self.pending:
            self._launch(a, b)
        self.pending = []

    def finish(self):
        self.flush()
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)


class DataParallel:
    """Wraps a :class:`mau_b200.UrbanPredictor`; the module itself (and therefore the reference's
    training loop, optimizer and checkpoint code) is used unchanged."""

    def __init__(self, model, group=None, bucket_mb: float = 8.0, broadcast_init: bool = True,
                 sync_bn: bool = False, grad_dtype: str = "fp32"):
        self.model, self.group = model, group
        self.net = model.model
        self.sync_bn = bool(sync_bn)
        self.grad_dtype = grad_dtype
        self.bucket_numel = int(bucket_mb * (1 << 20) / 4)
        object.__setattr__(self.net, "_dp", self)   # picked up by the model's forward -> HotPathFn
        if broadcast_init and dist.is_initialized() and dist.get_world_size(group) > 1:
            for t in model.state_dict().values():
                dist.broadcast(t, src=0, group=group)

    def prepare_plan(self, plan: "engine.Plan"):
        """Called by the model when it (re)uses a training plan: installs the SyncBN all-reduce if requested."""
        if not self.sync_bn or getattr(plan, "_sync_installed", False):
            return
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        group = self.group

        def allreduce(t: torch.Tensor):
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        plan.set_stats_sync(allreduce, world)
        plan._sync_installed = True

    # called by engine.HotPathFn.backward ---------------------------------------------------
    def make_grads(self, plan: "engine.Plan", diff_idx: List[int], shapes, params=None) -> Tuple[List[Optional[torch.Tensor]], list]:
        """Gradient tensors are views of one flat buffer laid out in state order, so a layer's
        (weight, bias, gamma, beta) -- contiguous state indices -- is one contiguous slice.

        Autograd may install such a view as ``param.grad`` without copying.  If a caller keeps gradients across
        backward passes (gradient accumulation, ``zero_grad(set_to_none=False)``), the next pass would overwrite
        them in place, so any ``.grad`` that still aliases the flat buffer is detached into its own storage first."""
        cache = getattr(plan, "_dp_cache", None)
        if cache is not None and params is not None:
            base = cache["flat"].untyped_storage().data_ptr()
            for p in params:
                if p.grad is not None and p.grad.untyped_storage().data_ptr() == base:
                    p.grad = p.grad.clone()
        if cache is None or cache["idx"] != list(diff_idx):
            offs, total = {}, 0
            for i, shp in zip(diff_idx, shapes):
                n = 1
                for s in shp:
                    n *= s
                offs[i] = (total, n, tuple(shp))
                total += n
            flat = torch.empty(total, device=plan.device, dtype=torch.float32)
            cache = {"idx": list(diff_idx), "offs": offs, "flat": flat,
                     "reducer": GradReducer(flat, self.group, self.bucket_numel, getattr(self, "grad_dtype", "fp32"))}
            plan._dp_cache = cache
            red = cache["reducer"]
            if hasattr(plan, "wait_backward_streams"):
                red.extra_wait = plan.wait_backward_streams

            def hook(first: int, last: int, offs=offs, red=red):
                lo = [offs[i][0] for i in range(first, last + 1) if i in offs]
                hi = [offs[i][0] + offs[i][1] for i in range(first, last + 1) if i in offs]
                if lo:
                    red.ready(min(lo), max(hi))
            plan.set_grad_hook(hook)
        grads_full: List[Optional[torch.Tensor]] = [None] * plan.num_state
        outs = []
        for i in diff_idx:
            a, n, shp = cache["offs"][i]
            g = cache["flat"][a:a + n].view(shp)
            grads_full[i] = g
            outs.append(g)
        return grads_full, outs

    def finish(self, plan: "engine.Plan"):
        plan._dp_cache["reducer"].finish()

    # ----------------------------------------------------------------------------------------
    def sync_buffers(self):
        """Average BatchNorm running statistics across ranks (call before saving a checkpoint)."""
        if not dist.is_initialized():
            return
        w = dist.get_world_size(self.group)
        for k, t in self.model.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                dist.all_reduce(t, group=self.group)
                t.div_(w)


def shard_tiles(n_tiles: int, rank: int, world: int) -> range:
    """Contiguous split of a tile list for inference (no collective needed)."""
    per = (n_tiles + world - 1) // world
    return range(min(n_tiles, rank * per), min(n_tiles, (rank + 1) * per))
