"""Host-side mirror of the reference's ``src/utils/losses.py`` on the device kernels of the hot path: same function
names, arguments and dictionary keys, every value a 0-d CUDA tensor, every entry differentiable w.r.t. ``outputs`` like
the reference's (a caller may recombine components with its own weights).

* pixel (L1 / MSE) and gradient-difference terms: ``loss_kernel`` (csrc/loss.cu), one pass forward + backward.
* SSIM term: ``ssim_*_kernel`` (csrc/ssim.cu).  The reference calls ``piq.ssim`` (src/utils/losses.py:3,88), a
  third-party package it does not pin and that is absent here; the kernels implement the algorithm piq publishes
  (11 x 11 Gaussian window, sigma 1.5, valid filtering, k1 = .01, k2 = .03).  Parity with piq itself is UNPINNED;
  the arithmetic is checked on the CPU against a torch restatement (oracle/ssim_oracle.py, tests/test_oracle.py).
  piq's input validation (values outside [0, 1] raise) is not reproduced.

Drop-in: ``from mau_b200.losses import compute_loss_mse, compute_loss_mse_gradient, compute_loss_l1_grad_ssim,
compute_all_loss`` in place of ``from src.utils.losses import ...`` (src/train.py:218-225, :38).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import engine


def gradient_loss(pred: torch.Tensor, target: torch.Tensor) -> Dict[str, torch.Tensor]:
    """src/utils/losses.py:5-25 (differentiable w.r.t. ``pred``)."""
    _, _, g = engine._LossFn.apply(pred, target, "l1", 0.0)
    return {"gradient": g}


def compute_loss_mse(outputs: torch.Tensor, targets: torch.Tensor) -> Dict[str, torch.Tensor]:
    """src/utils/losses.py:27-39."""
    _, mse, _ = engine._LossFn.apply(outputs, targets, "mse", 0.0)
    return {"total": mse, "mse": mse}


def compute_loss_mse_gradient(outputs, targets, lambda_grad: float = 0.1) -> Dict[str, torch.Tensor]:
    """src/utils/losses.py:41-57."""
    return engine.compute_loss_mse_gradient(outputs, targets, lambda_grad)


def compute_loss_l1_grad_ssim(outputs, targets, lambda_grad: float = 0.1, lambda_ssim: float = 0.5) -> Dict[str, torch.Tensor]:
    """src/utils/losses.py:59-99: L1 + lambda_grad * gradient + lambda_ssim * (1 - mean SSIM)."""
    d = engine.compute_loss_l1_grad(outputs, targets, lambda_grad)
    ssim = engine.ssim_loss(outputs, targets)
    return {"total": d["total"] + lambda_ssim * ssim, "pixel": d["pixel"], "gradient": d["gradient"], "ssim": ssim}


def compute_all_loss(outputs, targets, lambda_grad: float = 0.1, lambda_ssim: float = 0.5) -> Dict[str, torch.Tensor]:
    """src/utils/losses.py:101-115: the union of both dictionaries (the second overwrites 'total' and 'gradient')."""
    losses = {}
    losses.update(compute_loss_mse_gradient(outputs, targets, lambda_grad=lambda_grad))
    losses.update(compute_loss_l1_grad_ssim(outputs, targets, lambda_grad=lambda_grad, lambda_ssim=lambda_ssim))
    return losses
