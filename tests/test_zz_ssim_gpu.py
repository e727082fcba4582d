"""GPU: SSIM term of the training loss (csrc/ssim.cu, mau_b200.losses.compute_loss_l1_grad_ssim) against the torch
restatement oracle/ssim_oracle.py (parity with piq itself is unpinned: third party, not pinned by the reference, absent).

The arithmetic of the kernels (csrc/ssim_core.h) and the kernel source itself (csrc/ssim.cu compiled for the CPU on
oracle/cuda_emu.h) are also verified on the CPU against torch autograd (tests/test_oracle.py).  The check runs in its own
process (tools/ssim_gpu_check.py) so that a fault in these kernels cannot poison the CUDA context of the other GPU tests,
and the file sorts last.  It passed on the round-1 driver box (then still marked xfail); it is a hard requirement now."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_ssim_loss_kernels_against_torch_restatement():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ssim_gpu_check.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"], d
