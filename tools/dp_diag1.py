"""1-GPU diagnostic for tools/dp_parity.py: the single-process engine at the 8-rank GLOBAL batch (16 tiles of 37x45, fp32
mode) against the CPU oracle, with and without the second-stream overlap (MAU_FLAGS=8192), repeated to expose run-to-run
variation.  Prints median / worst per-tensor relative L2 error."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mau_b200  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
for mt, kw in (("unet", dict(temporal_embeddings=False, metadata_embeddings=True)), ("unet++", dict())):
    torch.manual_seed(7)
    m0 = mau_b200.UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
    sd0 = {k: v.clone() for k, v in m0.state_dict().items()}
    x, ts, md, tgt = O.synthetic_batch(16, 37, 45, T=24, seed=99)
    _, lref, grads, _ = O.train_step_grads(sd0, mt, x, ts, md, tgt, loss="mse", **kw)
    prev = None
    for rep in range(3):
        m1 = mau_b200.UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
        m1.load_state_dict(sd0)
        m1 = m1.to(dev).set_precision("fp32").train()
        out1 = m1(x.to(dev), ts.to(dev), md.to(dev))
        ((out1 - tgt.to(dev)) ** 2).mean().backward()
        torch.cuda.synchronize()
        g1 = {k: p.grad.cpu() for k, p in m1.named_parameters() if p.grad is not None}
        errs = sorted((float((g1[k] - grads[k]).norm() / grads[k].norm().clamp_min(1e-12)), k) for k in g1 if grads[k].norm() > 1e-7)
        line = f"{mt} flags={os.environ.get('MAU_FLAGS', '0')} rep {rep}: vs oracle median {errs[len(errs) // 2][0]:.2e} worst {errs[-1][0]:.2e} ({errs[-1][1]})"
        if prev is not None:
            d = sorted(float((g1[k] - prev[k]).norm() / prev[k].norm().clamp_min(1e-12)) for k in g1 if prev[k].norm() > 1e-7)
            line += f"; vs previous run median {d[len(d) // 2]:.2e} worst {d[-1]:.2e}"
        print(line, flush=True)
        prev = g1
