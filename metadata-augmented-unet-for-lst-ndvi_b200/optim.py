"""Fused AdamW for the training loop around the hot path.

``FusedAdamW`` is a drop-in for ``torch.optim.AdamW`` as the reference constructs it
(``src/train.py:213-214``: ``AdamW(model.parameters(), lr=..., weight_decay=...)``): same constructor
arguments, same ``param_groups`` / ``state`` layout (``step``, ``exp_avg``, ``exp_avg_sq`` per
parameter), so ``optimizer.state_dict()`` -- which the reference stores in its checkpoints
(``src/train.py:309``) -- is interchangeable with the stock optimizer's.  ``step()`` updates every
parameter that has a gradient with ONE launch of ``adamw_kernel`` (csrc/optim.cu) instead of several
element-wise kernels per tensor.  Parameters whose ``.grad`` is ``None`` (flag-disabled encoders) are
skipped exactly like the stock optimizer skips them: no weight decay, no state.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import engine


class FusedAdamW(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **kw):
        if kw.get("amsgrad") or kw.get("maximize"):
            raise ValueError("FusedAdamW: amsgrad / maximize are not supported")
        kw.pop("amsgrad", None); kw.pop("maximize", None)
        kw["foreach"], kw["fused"], kw["capturable"], kw["differentiable"] = False, False, False, False
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = engine.lib()
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if not g.is_contiguous() or g.dtype != torch.float32:
                    g = g.contiguous().float()
                st = self.state[p]
                if len(st) == 0:                      # same lazy state as torch.optim.AdamW
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault((int(st["step"]), p.device), []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            lr = float(group["lr"])
            b1, b2 = group["betas"]
            for (t, dev), items in by_step.items():
                n = len(items)
                arrs = [(C.c_void_p * n)(*[x[k].data_ptr() for x in items]) for k in range(4)]
                numels = (C.c_int64 * n)(*[x[0].numel() for x in items])
                with torch.cuda.device(dev):
                    engine.check(L.mau_adamw_step(n, arrs[0], arrs[1], arrs[2], arrs[3], numels, lr, float(b1), float(b2),
                                                  float(group["eps"]), float(group["weight_decay"]), t,
                                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "adamw_step")
        # the kernel wrote the parameters through raw pointers: torch's tensor._version did not move, so tell the
        # engine that every cached weight pack (eval plans) is stale
        engine.bump_state_epoch()
        return loss
