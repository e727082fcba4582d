"""SSIM loss kernels (csrc/ssim.cu) against the torch restatement oracle/ssim_oracle.py on cuda:0.
Prints one JSON line {"ok": bool, "cases": [...]}.  Run in its own process by tests/test_zz_ssim_gpu.py so that a
faulting kernel cannot poison the CUDA context of the main test process."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mau_b200  # noqa: E402
from mau_b200 import engine, losses  # noqa: E402
from oracle import ssim_oracle as S  # noqa: E402


def main():
    cases, ok = [], True
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 2, 23, 31), (1, 2, 11, 11), (3, 2, 40, 17), (2, 3, 30, 30), (16, 2, 250, 250), (2, 2, 512, 512), (1, 2, 385, 400)):
        out = torch.randn(shape, generator=g) * 0.7
        tgt = torch.randn(shape, generator=g) * 0.7
        tgt[:, 0] = tgt[:, 0].clamp(-1, 1)
        ref_in = out.clone().requires_grad_(True)
        ref = S.compute_loss_l1_grad_ssim(ref_in, tgt)
        ref["total"].backward()
        dev_in = out.cuda().requires_grad_(True)
        got = losses.compute_loss_l1_grad_ssim(dev_in, tgt.cuda())
        got["total"].backward()
        torch.cuda.synchronize()
        e_ssim = abs(float(got["ssim"]) - float(ref["ssim"]))
        e_total = abs(float(got["total"]) - float(ref["total"]))
        gref = ref_in.grad
        e_grad = float((dev_in.grad.cpu() - gref).abs().max() / gref.abs().max())
        # SSIM gradient alone
        dev2 = out.cuda().requires_grad_(True)
        engine.ssim_loss(dev2, tgt.cuda()).backward()
        ref2 = out.clone().requires_grad_(True)
        S.ssim_loss(ref2, tgt).backward()
        e_sgrad = float((dev2.grad.cpu() - ref2.grad).abs().max() / ref2.grad.abs().max())
        good = e_ssim < 2e-5 and e_total < 5e-5 and e_grad < 1e-3 and e_sgrad < 1e-3
        ok &= good
        cases.append({"shape": list(shape), "ssim_abs_err": e_ssim, "total_abs_err": e_total, "grad_rel_err": e_grad,
                      "ssim_grad_rel_err": e_sgrad, "ok": good})
    with torch.no_grad():      # validation path (src/train.py:33-42): values only, no backward kernel
        val = losses.compute_all_loss(out.cuda(), tgt.cuda())
    ok &= abs(float(val["ssim"]) - float(ref["ssim"])) < 2e-5 and set(val) == {"total", "mse", "gradient", "pixel", "ssim"}
    x = torch.rand(2, 2, 32, 32, generator=g)
    same = float(engine.ssim_loss(x.cuda(), x.cuda()))      # identical maps -> SSIM 1 -> loss 0
    ok &= abs(same) < 1e-6
    print(json.dumps({"ok": bool(ok), "identical_maps_loss": same, "cases": cases}))


if __name__ == "__main__":
    main()
