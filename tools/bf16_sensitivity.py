#!/usr/bin/env python
"""How tight can a whole-model bf16 TRAINING comparison be?  (CPU only, no GPU needed.)

Runs one training step (train-mode BatchNorm, MSE loss) of the oracle three ways on the same inputs:
  A  fp32 everywhere (the pinned oracle)
  B  bf16 rounding at the engine's storage points, fp32 arithmetic between stores  (oracle emulate_bf16=True)
  C  the same storage points, fp64 arithmetic between stores
B and C are two equally valid "bf16 engines": they round the same tensors at the same places and differ only in
the last bits of the values that get rounded (what a different summation order does on a GPU).  The table shows
how far apart they end up -- the floor for any whole-model bf16-vs-oracle tolerance -- next to the bf16-vs-fp32 gap.
Per-tensor numbers: relative L2 over every parameter gradient except conv biases in front of BatchNorm (zero).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mau_b200  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402


def run(mt, kw, ctor, bf, B, H, W, T, seed):
    torch.manual_seed(seed)
    m = mau_b200.UrbanPredictor(mt, *ctor, base_filters=bf, **kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x, ts, md, tgt = O.synthetic_batch(B, H, W, T=T, seed=1003)
    oa, la, ga, _ = O.train_step_grads(sd, mt, x, ts, md, tgt, loss="mse", **kw)
    ob, lb, gb, _ = O.train_step_grads(sd, mt, x, ts, md, tgt, loss="mse", emulate_bf16=True, **kw)
    keep = O._bf16
    O._bf16 = lambda t: t.to(torch.float32).to(torch.bfloat16).to(t.dtype)
    try:
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        oc, lc, gc, _ = O.train_step_grads(sd64, mt, x.double(), ts.double(), md.double(), tgt.double(), loss="mse",
                                           emulate_bf16=True, **kw)
    finally:
        O._bf16 = keep

    def stats(g1, g2):
        e = []
        for k in g1:
            if g1[k] is None or (".conv" in k and k.endswith(".bias")):
                continue
            e.append(float((g1[k].double() - g2[k].double()).norm() / max(float(g2[k].double().norm()), 1e-30)))
        e.sort()
        return e[len(e) // 2], e[-1]

    def orel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())
    return dict(out_ba=orel(ob, oa), out_bc=orel(ob, oc), g_ba=stats(gb, ga), g_bc=stats(gb, gc))


if __name__ == "__main__":
    kw = dict(temporal_embeddings=False, metadata_embeddings=True)
    small, full = (23, 828, 16, 8, 8, 32, 2), (23, 828, 64, 8, 64, 96, 2)
    cases = [("U-Net 8 filters, 3x37x45", "unet", kw, small, 8, 3, 37, 45, 24, 123),
             ("U-Net++ 8 filters, 3x37x45", "unet++", {}, small, 8, 3, 37, 45, 24, 123),
             ("U-Net 64 filters, 4x64x64", "unet", kw, full, 64, 4, 64, 64, 8, 42),
             ("U-Net 64 filters, 4x128x128", "unet", kw, full, 64, 4, 128, 128, 8, 42)]
    if "--full" in sys.argv:
        cases.append(("U-Net 64 filters, 2x250x250", "unet", kw, full, 64, 2, 250, 250, 8, 42))
    print("| case | train-mode output, bf16 vs fp32 | output, bf16(fp32 arith) vs bf16(fp64 arith) | grads bf16 vs fp32: median / worst |"
          " grads bf16 vs bf16: median / worst |")
    print("|---|---:|---:|---:|---:|")
    for name, mt, k, ctor, bf, B, H, W, T, seed in cases:
        r = run(mt, k, ctor, bf, B, H, W, T, seed)
        print(f"| {name} | {r['out_ba']:.1e} | {r['out_bc']:.1e} | {r['g_ba'][0]:.1e} / {r['g_ba'][1]:.1e} | "
              f"{r['g_bc'][0]:.1e} / {r['g_bc'][1]:.1e} |", flush=True)
