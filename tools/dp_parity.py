"""N-rank data-parallel training step vs the single-process oracle at the GLOBAL batch.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_parity.py

With sync_bn=True the step must equal the reference semantics (one device, whole batch): loss, every
parameter gradient (after the gradient average) and the BatchNorm running statistics.  fp32 mode.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import mau_b200  # noqa: E402
from mau_b200 import parallel  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grad_dtype = "bf16" if "--grad-dtype=bf16" in sys.argv else "fp32"      # wire format of the gradient all-reduce
    tol1 = 1e-3 if grad_dtype == "fp32" else 1e-2                            # bf16 buckets: every gradient rounded twice
    ok_all = True
    for mt, kw in (("unet", dict(temporal_embeddings=False, metadata_embeddings=True)), ("unet++", dict())):
        torch.manual_seed(7)
        m = mau_b200.UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        per = 2
        x, ts, md, tgt = O.synthetic_batch(per * world, 37, 45, T=24, seed=99)
        m = m.to(dev).set_precision("fp32").train()
        parallel.DataParallel(m, sync_bn=True, grad_dtype=grad_dtype)
        sl = slice(rank * per, (rank + 1) * per)
        out = m(x[sl].to(dev), ts[sl].to(dev), md[sl].to(dev))
        loss = ((out - tgt[sl].to(dev)) ** 2).mean()      # MSE: an L1 loss has a sign() gradient (ill-conditioned parity)
        loss.backward()
        torch.cuda.synchronize()
        lsum = loss.detach().clone()
        dist.all_reduce(lsum)
        if rank == 0:
            # (1) the same engine, one process, the GLOBAL batch: differs from the N-rank run only by the order of the
            #     SyncBN / gradient reductions -> tight tolerance
            torch.manual_seed(7)
            m1 = mau_b200.UrbanPredictor(mt, 23, 828, 16, 8, 8, 32, 2, base_filters=8, **kw)
            m1.load_state_dict(sd0)
            m1 = m1.to(dev).set_precision("fp32").train()
            out1 = m1(x.to(dev), ts.to(dev), md.to(dev))
            ((out1 - tgt.to(dev)) ** 2).mean().backward()
            torch.cuda.synchronize()
            g1 = {k: p.grad for k, p in m1.named_parameters() if p.grad is not None}
            # (2) the CPU oracle at the global batch (train-mode BatchNorm makes single tiny-norm tensors ill-conditioned
            #     in fp32 -- ReLU mask flips -- hence the absolute floor, as in tests/test_gpu_parity.py)
            _, lref, grads, sd1 = O.train_step_grads(sd0, mt, x, ts, md, tgt, loss="mse", **kw)
            worst1, name1, worst2, name2 = 0.0, "", 0.0, ""
            all1 = []
            ok = True
            for k, p in m.named_parameters():
                if p.grad is None:
                    continue
                e1 = float((p.grad - g1[k]).norm() / g1[k].norm().clamp_min(1e-12))
                if g1[k].norm() > 1e-7:
                    all1.append(e1)
                if g1[k].norm() > 1e-7 and e1 > worst1:
                    worst1, name1 = e1, k
                r = grads[k]
                err2, gn = float((p.grad.cpu() - r).norm()), float(r.norm())
                ok &= err2 <= (2e-2 if grad_dtype == "fp32" else 3e-2) * gn + 2e-6
                if gn > 1e-6 and err2 / gn > worst2:
                    worst2, name2 = err2 / gn, k
            sd_now = m.state_dict()
            rs = max(float((sd_now[k].cpu() - sd1[k]).abs().max() / sd1[k].abs().max().clamp_min(1.0))
                     for k in sd1 if "running" in k)
            lerr = abs(float(lsum) / world - float(lref)) / max(abs(float(lref)), 1e-12)
            # The N-rank run and the single-process run build the BatchNorm sums from fp32 partials over different pixel
            # groupings (2 tiles per rank vs 16 in one process): mean / rstd differ in the last fp32 bit, an activation within
            # rounding of zero can then fall on the other side of the ReLU, and ONE such element moves its channel's dbeta by
            # ~1/sqrt(pixels) and -- when it sits in the last decoder block -- every upstream gradient of this 8-filter model
            # by ~1e-3.  The same mechanism separates the single-process engine from the CPU oracle at this batch (median 3e-4,
            # bit-reproducible run to run and identical with the stream overlap on or off: tools/dp_diag1.py).  Forward
            # quantities (loss, running statistics) stay at rounding level, and at 2 ranks (4 tiles) no element flips: 5e-6.
            all1.sort()
            med1 = all1[len(all1) // 2] if all1 else 0.0
            flip_floor = 2e-3 if world * per > 4 else 2e-5
            ok = ok and med1 < max(flip_floor, tol1 if grad_dtype != "fp32" else 0.0) and worst1 < max(tol1, 5 * flip_floor) \
                and rs < 1e-4 and lerr < 1e-5
            ok_all &= ok
            print(f"[dp_parity] {mt} world={world} grads on the wire: {grad_dtype}: loss err {lerr:.2e}; vs single-process engine at the global batch: grad rel L2 median {med1:.2e}, worst "
                  f"{worst1:.2e} ({name1}); vs CPU oracle: worst {worst2:.2e} ({name2}), running-stat err {rs:.2e} -> "
                  f"{'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
