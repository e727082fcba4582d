"""Helper of tests/test_oracle.py (run as a script, in its own process): executes tools/ssim_gpu_check.py WITHOUT a GPU.
``Tensor.cuda()`` becomes a detached copy, the SSIM entry point runs csrc/ssim.cu under oracle/cuda_emu.h (library path
in argv[1]), the L1 + gradient-difference kernel is replaced by torch ops.  Everything between the check script and the
kernels -- mau_b200.losses, the autograd Functions of engine.py, the JSON the GPU test parses -- is thereby exercised
before the first hardware run."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mau_b200  # noqa: E402,F401
from mau_b200 import engine  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self.detach().clone()
torch.cuda.synchronize = lambda *a, **k: None
L = C.CDLL(sys.argv[1])
L.emu_ssim_work_floats.restype = C.c_longlong
L.emu_ssim_loss.argtypes = [C.c_void_p] * 2 + [C.c_int] * 4 + [C.c_void_p] * 4


def emulated_ssim_terms(pred, target, need_grad=True):
    p, t = pred.detach().contiguous().float().numpy(), target.detach().contiguous().float().numpy()
    B, Cc, H, W = p.shape
    work = np.zeros(max(L.emu_ssim_work_floats(B, H, W), 1), np.float32)
    acc, lv, g = np.zeros(1), np.zeros(1, np.float32), np.zeros_like(p)
    rc = L.emu_ssim_loss(p.ctypes.data, t.ctypes.data, B, Cc, H, W, lv.ctypes.data, g.ctypes.data if need_grad else None,
                         work.ctypes.data, acc.ctypes.data)
    if rc:
        raise RuntimeError("emulated ssim_loss failed")
    return torch.from_numpy(lv), (torch.from_numpy(g) if need_grad else None)


def torch_loss_terms(pred, target, kind="l1", lambda_grad=0.1, need_grad=True):
    with torch.enable_grad():
        x = pred.detach().clone().requires_grad_(True)
        pix = F.l1_loss(x, target) if kind == "l1" else F.mse_loss(x, target)
        dy = ((x[:, :, 1:] - x[:, :, :-1]).abs() - (target[:, :, 1:] - target[:, :, :-1]).abs()).abs().mean()
        dx = ((x[:, :, :, 1:] - x[:, :, :, :-1]).abs() - (target[:, :, :, 1:] - target[:, :, :, :-1]).abs()).abs().mean()
        tot = pix + lambda_grad * (dy + dx)
        tot.backward()
    return torch.stack([tot.detach(), pix.detach(), (dy + dx).detach(), torch.zeros(())]), x.grad


def torch_loss_backward_terms(pred, target, kind, lambda_total, g_total=None, g_pixel=None, g_grad=None):
    with torch.enable_grad():
        x = pred.detach().clone().requires_grad_(True)
        pix = F.l1_loss(x, target) if kind == "l1" else F.mse_loss(x, target)
        dy = ((x[:, :, 1:] - x[:, :, :-1]).abs() - (target[:, :, 1:] - target[:, :, :-1]).abs()).abs().mean()
        dx = ((x[:, :, :, 1:] - x[:, :, :, :-1]).abs() - (target[:, :, :, 1:] - target[:, :, :, :-1]).abs()).abs().mean()
        tot = pix + lambda_total * (dy + dx)
        z = torch.zeros(())
        ((g_total if g_total is not None else z) * tot + (g_pixel if g_pixel is not None else z) * pix
         + (g_grad if g_grad is not None else z) * (dy + dx)).backward()
    return x.grad


def emulated_ssim_forward(pred, target):
    lv, _ = emulated_ssim_terms(pred, target, need_grad=False)
    return lv, None


def emulated_ssim_backward(pred, target, work, upstream):
    _, g = emulated_ssim_terms(pred, target, need_grad=True)
    return g if upstream is None else g * upstream


engine.ssim_loss_terms = emulated_ssim_terms
engine.ssim_forward_terms = emulated_ssim_forward
engine.ssim_backward_terms = emulated_ssim_backward
engine.loss_terms = torch_loss_terms
engine.loss_backward_terms = torch_loss_backward_terms
src = open(os.path.join(ROOT, "tools", "ssim_gpu_check.py")).read()
assert "(16, 2, 250, 250)" in src
src = src.replace(", (2, 2, 512, 512), (1, 2, 385, 400)", "")      # emulated blocks are real threads: keep it small
exec(compile(src.replace("(16, 2, 250, 250)", "(1, 2, 40, 60)"), "ssim_gpu_check.py", "exec"))
