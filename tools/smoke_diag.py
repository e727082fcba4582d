"""Which path moves the bf16 training gradient of the smoke() configuration? (debug helper)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mau_b200
from oracle import unet_oracle as O
kw = dict(temporal_embeddings=False, metadata_embeddings=True)
def run(precision, label):
    torch.manual_seed(123)
    m = mau_b200.UrbanPredictor("unet", 23, 828, 16, 8, 16, 32, 2, base_filters=16, **kw)
    O.perturb_bn_stats(m.state_dict())
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x, ts, md, tgt = O.synthetic_batch(2, 50, 50, T=40, seed=1002)
    m = m.to("cuda:0").set_precision(precision).train()
    out = m(x.cuda(), ts.cuda(), md.cuda())
    loss = (out - tgt.cuda()).abs().mean()
    loss.backward()
    torch.cuda.synchronize()
    oref, lref, grads, _ = O.train_step_grads(sd, "unet", x, ts, md, tgt, loss="l1", **kw)
    errs = {}
    for k, p in m.named_parameters():
        if p.grad is not None and grads[k].norm() > 1e-7:
            errs[k] = float((p.grad.cpu() - grads[k]).norm() / grads[k].norm())
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    oe = float((out.detach().cpu() - oref).abs().max() / oref.abs().max())
    print(f"{label:28s} out err {oe:.3e} loss {float(loss):.5f}/{float(lref):.5f} conv0_1.conv2.w {errs['model.conv0_1.conv2.weight']:.3e} worst {[(k[6:], round(v, 3)) for k, v in worst]}", flush=True)
run("fp32", "fp32")
run("bf16", "bf16 " + os.environ.get("TAG", "default"))
