"""Generate tests/golden/* from the REAL reference (runs only in the build container,
where /root/reference is mounted; the GPU box never runs this).

    python oracle/gen_golden.py

Recipe: torch.manual_seed(42); build UrbanPredictor(variant); inputs from
torch.Generator().manual_seed(7): x=randn(2,23,50,50), ts=randn(2,T), md=randn(2,8);
eval forward; then train forward, loss = y.abs().mean(), backward.  H=50 exercises the
two-stage bilinear path (12 -> 24 -> 25).
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

VARIANTS = {
    # name: (model_type, T, kwargs)
    "unet_noemb": ("unet", 828, dict(temporal_embeddings=False, metadata_embeddings=False)),
    "unet_metaemb": ("unet", 828, dict(temporal_embeddings=False, metadata_embeddings=True)),
    "unet_emb": ("unet", 60, dict(temporal_embeddings=True, metadata_embeddings=True)),
    "unetpp_emb": ("unet++", 60, dict()),
}
CTOR = (23, 828, 64, 8, 64, 96, 2)


def main():
    sys.path.insert(0, REF)
    from src.model import UrbanPredictor  # the real reference

    os.makedirs(OUT, exist_ok=True)
    keys = {}
    for name, (mt, T, kw) in VARIANTS.items():
        torch.manual_seed(42)
        m = UrbanPredictor(mt, *CTOR, **kw)
        sd = m.state_dict()
        keys[name] = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
        g = torch.Generator().manual_seed(7)
        x = torch.randn(2, 23, 50, 50, generator=g)
        ts = torch.randn(2, T, generator=g)
        md = torch.randn(2, 8, generator=g)
        m.eval()
        with torch.no_grad():
            y_eval = m(x, ts, md)
        m.train()
        y_tr = m(x, ts, md)
        loss = y_tr.abs().mean()
        loss.backward()
        gn = {n: (float(p.grad.norm()) if p.grad is not None else None) for n, p in m.named_parameters()}
        sd2 = m.state_dict()
        # a few full gradient tensors (small ones) for element-wise checks
        gsmall = {n: p.grad.numpy() for n, p in m.named_parameters()
                  if p.grad is not None and p.numel() <= 4096}
        np.savez_compressed(
            os.path.join(OUT, f"kat_{name}.npz"),
            y_eval=y_eval.numpy(), y_train=y_tr.detach().numpy(), loss=np.float32(loss.item()),
            grad_l2=np.float32(torch.sqrt(sum(p.grad.pow(2).sum() for p in m.parameters()
                                              if p.grad is not None)).item()),
            bn_rm=sd2["model.conv0_0.bn1.running_mean"].numpy(),
            bn_rv=sd2["model.conv0_0.bn1.running_var"].numpy(),
            bn_count=sd2["model.conv0_0.bn1.num_batches_tracked"].numpy(),
            grad_norms=json.dumps(gn),
            **{"g::" + k: v for k, v in gsmall.items()})
        print(name, float(y_eval.sum()), float(y_eval.abs().mean()), float(loss))

    # small-filter variants (fast everywhere): full outputs for odd sizes incl. 2-stage resize
    for name, (mt, T, kw) in VARIANTS.items():
        torch.manual_seed(123)
        m = UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(3, 23, 37, 45, generator=g)
        ts = torch.randn(3, 40, generator=g)
        md = torch.randn(3, 8, generator=g)
        m.eval()
        with torch.no_grad():
            y = m(x, ts, md)
        np.savez_compressed(os.path.join(OUT, f"small_{name}.npz"), y_eval=y.numpy())
    # deep supervision (never enabled by reference callers, kept for completeness)
    torch.manual_seed(123)
    m = UrbanPredictor("unet++", 23, 828, 16, 8, 8, 32, 2, base_filters=8, deep_supervision=True)
    keys["unetpp_ds_small"] = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 23, 37, 45, generator=g); ts = torch.randn(3, 40, generator=g); md = torch.randn(3, 8, generator=g)
    m.eval()
    with torch.no_grad():
        ys = m(x, ts, md)
    np.savez_compressed(os.path.join(OUT, "small_unetpp_ds.npz"), y_eval=np.stack([t.numpy() for t in ys]))

    with open(os.path.join(OUT, "state_keys.json"), "w") as f:
        json.dump(keys, f)

    # loss terms from the reference's own src/utils/losses.py (piq stubbed: SSIM not pinned)
    stub = types.ModuleType("piq")
    stub.ssim = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("piq not available"))
    sys.modules["piq"] = stub
    from src.utils.losses import gradient_loss, compute_loss_mse_gradient
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    pred = torch.randn(3, 2, 33, 47, generator=g, requires_grad=True)
    tgt = torch.randn(3, 2, 33, 47, generator=g)
    gl = gradient_loss(pred, tgt)["gradient"]
    l1 = F.l1_loss(pred, tgt)                       # src/utils/losses.py:67
    tot = l1 + 0.1 * gl
    (gr,) = torch.autograd.grad(tot, pred)
    mg = compute_loss_mse_gradient(pred, tgt, 0.1)
    (gr2,) = torch.autograd.grad(mg["total"], pred)
    np.savez_compressed(os.path.join(OUT, "loss_terms.npz"), pred=pred.detach().numpy(), tgt=tgt.numpy(),
                        l1=l1.item(), grad=gl.item(), total_l1=tot.item(), dpred_l1=gr.numpy(),
                        mse=mg["mse"].item(), total_mse=mg["total"].item(), dpred_mse=gr2.numpy())
    print("golden written to", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
