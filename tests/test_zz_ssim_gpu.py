"""GPU: SSIM term of the training loss (csrc/ssim.cu, mau_b200.losses.compute_loss_l1_grad_ssim) against the torch
restatement oracle/ssim_oracle.py (parity with piq itself is unpinned: third party, not pinned by the reference, absent).

The arithmetic of the kernels (csrc/ssim_core.h) and the kernel source itself (csrc/ssim.cu compiled for the CPU on
oracle/cuda_emu.h) ARE verified on the CPU against torch autograd (tests/test_oracle.py); the kernels were written after
the round's GPU budget was spent and have not run on hardware yet, hence the non-strict xfail: the check runs in its own process (tools/ssim_gpu_check.py) so that a fault in the new kernels cannot
poison the CUDA context of the other GPU tests, and the file sorts last."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="SSIM kernels not yet run on hardware (round-1 GPU budget spent); verified on the CPU under emulation")
def test_ssim_loss_kernels_against_torch_restatement():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ssim_gpu_check.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"], d
