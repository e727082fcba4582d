#!/usr/bin/env python
"""Kernel timeline of one training step (config 3) from CUPTI via torch.profiler: which stream runs what, how much the
streams overlap, where the device idles.  Writes gpurun_out/<tag>_timeline.json (one record per kernel of one step) and
prints a summary.  nsys is not installed in the image; this is the timeline evidence for the backward overlap and for
the NCCL kernels of the data-parallel step (run under torchrun for N > 1: rank 0 reports).

    python tools/timeline.py [--config 3] [--tag r02] [--batch 16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mau_b200  # noqa: E402
from mau_b200 import engine  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

CTOR = (23, 828, 64, 8, 64, 96, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--grad-dtype", default="fp32")
    ap.add_argument("--comm-ctas", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas, opts.config.min_ctas = args.comm_ctas, 1
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        engine.lib().mau_set_sm_reserve(args.comm_ctas)
    mt, kw = ("unet++", {}) if args.config == 4 else ("unet", dict(temporal_embeddings=False, metadata_embeddings=True))
    torch.manual_seed(42)
    model = mau_b200.UrbanPredictor(mt, *CTOR, **kw).to(dev).train()
    if world > 1:
        from mau_b200 import parallel
        parallel.DataParallel(model, grad_dtype=args.grad_dtype)
    opt = mau_b200.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-3)
    batches = [[t.to(dev) for t in O.synthetic_batch(args.batch, 250, 250, seed=1002 + i + 17 * rank)] for i in range(2)]

    def step(i):
        x, ts, md, tgt = batches[i % 2]
        loss = engine.compute_loss_l1_grad(model(x, ts, md), tgt, 0.0)["total"]
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    recs = []
    for e in ev:
        recs.append({"name": e.name, "start_us": e.time_range.start, "dur_us": e.time_range.end - e.time_range.start,
                     "stream": getattr(e, "stream", None) if hasattr(e, "stream") else None})
    recs.sort(key=lambda r: r["start_us"])
    # the middle step: between the 2nd and the 3rd nchw_to_nhwc launch
    starts = [i for i, r in enumerate(recs) if "nchw_to_nhwc" in r["name"]]
    a, b = (starts[1], starts[2]) if len(starts) >= 3 else (0, len(recs))
    one = recs[a:b]
    t0 = one[0]["start_us"]
    for r in one:
        r["start_us"] -= t0
    span = max(r["start_us"] + r["dur_us"] for r in one)
    # union busy time / idle gaps
    iv = sorted((r["start_us"], r["start_us"] + r["dur_us"]) for r in one)
    busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    tot = sum(r["dur_us"] for r in one)

    def fam(n):
        for k in ("wgrad3x3", "conv3x3", "bn_", "bilinear", "maxpool", "nccl", "adamw", "head", "pack_w", "wgrad_finalize", "loss", "cast_"):
            if k in n:
                return k
        return "other"
    by = {}
    for r in one:
        by.setdefault(fam(r["name"]), [0, 0.0])
        by[fam(r["name"])][0] += 1
        by[fam(r["name"])][1] += r["dur_us"]
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        out = os.path.join(ROOT, "gpurun_out", f"{args.tag}_timeline_c{args.config}_n{world}.json")
        json.dump(one, open(out, "w"))
        print(f"step span {span:.0f} us, {len(one)} kernels, summed kernel time {tot:.0f} us, union busy {busy:.0f} us, "
              f"idle {span - busy:.0f} us, overlapped {tot - busy:.0f} us")
        for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            print(f"  {k:16s} {n:4d} launches {t:8.0f} us")
        print("largest idle gaps (us @ time):", [(round(g, 1), round(t)) for g, t in sorted(gaps, reverse=True)[:8]])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
