"""LSTM gradient parity vs the oracle for several batch sizes / T (single GPU, fp32 mode)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mau_b200
from oracle import unet_oracle as O
for B, T, seed in [(2, 24, 99), (4, 24, 99), (2, 24, 7), (3, 40, 5), (2, 60, 3)]:
    torch.manual_seed(7)
    m = mau_b200.UrbanPredictor("unet++", 23, 828, 16, 8, 8, 32, 2, base_filters=8)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x, ts, md, tgt = O.synthetic_batch(B, 37, 45, T=T, seed=seed)
    m = m.cuda().set_precision("fp32").train()
    out = m(x.cuda(), ts.cuda(), md.cuda())
    loss = ((out - tgt.cuda()) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    _, lref, grads, _ = O.train_step_grads(sd0, "unet++", x, ts, md, tgt, loss="mse")
    res = []
    for k, p in m.named_parameters():
        if "temporal_encoder" in k or "meta_encoder" in k:
            r = grads[k]
            res.append((k.split("model.")[1], float((p.grad.cpu() - r).norm() / r.norm().clamp_min(1e-20)), float(r.norm())))
    print(f"B={B} T={T} seed={seed}: " + "  ".join(f"{n}:{e:.1e}(|g|={g:.1e})" for n, e, g in res), flush=True)
