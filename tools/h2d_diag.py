#!/usr/bin/env python
"""Concurrent host->device copy bandwidth per GPU at N ranks (names the limiter of the inference e2e number at N >= 4:
each rank pushes 92 MB of pinned fp32 NCHW per 1.8 ms step).  Run under torchrun; every rank copies the batch-16 input
block (pinned) to its GPU back to back for ~1 s after a barrier; rank 0 prints min / mean / sum GB/s.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_diag.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 16 * 23 * 250 * 250                       # one inference step's maps, fp32
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
    for h in host:
        h.normal_()
    dst = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3)]
    res = {}
    for dtype_name, div in (("fp32 NCHW (the reference's contract)", 1), ("bf16 (half the bytes)", 2)):
        m = n // div
        for _ in range(3):
            for h, d in zip(host, dst):
                d[:m].copy_(h[:m], non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        reps = 0
        while time.perf_counter() - t0 < 1.0:
            for h, d in zip(host, dst):
                d[:m].copy_(h[:m], non_blocking=True)
            reps += 3
            torch.cuda.current_stream().synchronize() if reps % 30 == 0 else None
        e1.record()
        torch.cuda.synchronize()
        gbs = reps * m * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9
        t = torch.tensor([gbs], device=dev)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v) for v in allv]
        else:
            vals = [gbs]
        res[dtype_name] = {"per_gpu_gbs_min": min(vals), "per_gpu_gbs_mean": sum(vals) / len(vals), "sum_gbs": sum(vals),
                           "bytes_per_copy": m * 4}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "cpu_count": os.cpu_count(), "h2d": res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
