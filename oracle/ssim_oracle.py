"""CPU oracle for the SSIM term of the reference's training loss (src/utils/losses.py:72-95).

TEST INFRASTRUCTURE ONLY (tests/ and tools/ import it; nothing in the product package does).

PARITY UNPINNED.  The reference calls ``piq.ssim`` (requirements.txt:9, no version pin, not vendored, absent from
the build image and from the GPU boxes), and no reference test holds a known answer for it.  This file restates the
algorithm piq publishes for ``ssim(x, y, kernel_size=11, kernel_sigma=1.5, data_range=1., reduction='none',
downsample=True, k1=0.01, k2=0.03)`` with plain torch ops so that autograd supplies the gradient:

  * inputs divided by data_range; average-pooled by f = max(1, round(min(H, W) / 256)) when f > 1;
  * window = outer product of exp(-(i - 5)^2 / (2 sigma^2)), normalised to sum 1, one copy per channel;
  * mu_x = conv2d(x, window, padding=0, groups=C), likewise mu_y, E[x^2], E[y^2], E[xy];
  * cs = (2 s_xy + c2) / (s_xx + s_yy + c2);  ss = (2 mu_x mu_y + c1) / (mu_x^2 + mu_y^2 + c1) * cs;
  * per-image value = mean over the valid positions, then over channels.

What IS pinned, on the CPU, is the product's own arithmetic against this restatement: tests/test_oracle.py compiles
oracle/ssim_host.cpp around csrc/ssim_core.h (the header the CUDA kernels include) and compares value and gradient
with this file's autograd.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gaussian_window(kernel_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    coords = torch.arange(kernel_size, dtype=torch.float32) - (kernel_size - 1) / 2.0
    g = coords ** 2
    g = (-(g.unsqueeze(0) + g.unsqueeze(1)) / (2 * sigma ** 2)).exp()
    return (g / g.sum()).unsqueeze(0)


def ssim(x: torch.Tensor, y: torch.Tensor, kernel_size: int = 11, kernel_sigma: float = 1.5, data_range: float = 1.0,
         k1: float = 0.01, k2: float = 0.03, downsample: bool = True, force_pool: int = 0) -> torch.Tensor:
    """Per-image SSIM, [B] (piq.ssim(..., reduction='none')).  ``force_pool`` (tests only) replaces piq's factor
    f = max(1, round(min(H, W) / 256)) so that the pooled path can be exercised on small tiles."""
    x, y = x / float(data_range), y / float(data_range)
    f = force_pool if force_pool > 0 else max(1, round(min(x.shape[-2:]) / 256))
    if f > 1 and downsample:
        x, y = F.avg_pool2d(x, kernel_size=f), F.avg_pool2d(y, kernel_size=f)
    if x.shape[-1] < kernel_size or x.shape[-2] < kernel_size:
        raise ValueError(f"Kernel size can't be greater than actual input size. Input size: {tuple(x.shape)}. Kernel size: {kernel_size}")
    C = x.shape[1]
    kernel = gaussian_window(kernel_size, kernel_sigma).to(x).repeat(C, 1, 1, 1)
    c1, c2 = k1 ** 2, k2 ** 2
    mu_x = F.conv2d(x, kernel, groups=C)
    mu_y = F.conv2d(y, kernel, groups=C)
    mu_xx, mu_yy, mu_xy = mu_x ** 2, mu_y ** 2, mu_x * mu_y
    s_xx = F.conv2d(x ** 2, kernel, groups=C) - mu_xx
    s_yy = F.conv2d(y ** 2, kernel, groups=C) - mu_yy
    s_xy = F.conv2d(x * y, kernel, groups=C) - mu_xy
    cs = (2.0 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2.0 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
    return ss.mean(dim=(-1, -2)).mean(1)


def ssim_loss(outputs: torch.Tensor, targets: torch.Tensor, force_pool: int = 0) -> torch.Tensor:
    """``1 - mean(ssim(scaled outputs, scaled targets))`` exactly as src/utils/losses.py:72-90 builds it."""
    t = torch.stack([(targets[:, 0] + 1.0) / 2.0, torch.clamp(targets[:, 1], 0.0, 1.0)], dim=1)
    o = torch.stack([(outputs[:, 0] + 1.0) / 2.0, torch.clamp(outputs[:, 1], 0.0, 1.0)], dim=1)
    return 1 - ssim(o, t, data_range=1.0, force_pool=force_pool).mean()


def compute_loss_l1_grad_ssim(outputs, targets, lambda_grad=0.1, lambda_ssim=0.5):
    """src/utils/losses.py:59-99 with the restated SSIM."""
    pixel = F.l1_loss(outputs, targets)
    dy_p = (outputs[:, :, 1:, :] - outputs[:, :, :-1, :]).abs()
    dx_p = (outputs[:, :, :, 1:] - outputs[:, :, :, :-1]).abs()
    dy_t = (targets[:, :, 1:, :] - targets[:, :, :-1, :]).abs()
    dx_t = (targets[:, :, :, 1:] - targets[:, :, :, :-1]).abs()
    grad = (dy_p - dy_t).abs().mean() + (dx_p - dx_t).abs().mean()
    s = ssim_loss(outputs, targets)
    return {"total": pixel + lambda_grad * grad + lambda_ssim * s, "pixel": pixel, "gradient": grad, "ssim": s}
