"""Input-pipeline benchmark: tiles/s of the native tile reader (mau_b200.data) against the
reference's loading path (oracle/dataset_oracle.py: np.load per sample on one thread + stack + pad, which is
what src/dataset.py does with num_workers=0) on the same synthetic archives of the SURVEY 8d shape.

    python tools/loader_bench.py [--tiles 64] [--edge 250] [--batch 16] [--threads 1,2,4,8] [--dir /dev/shm/...]
                                 [--device cpu|cuda:0]

Prints one JSON line.  `archive_MBps` is compressed bytes consumed per second, `payload_GBps` decoded fp32 bytes
per second.  With --device cuda:0 the batches also go host->device through the loader's pinned ring (the timed
region then ends with a device synchronize)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mau_b200  # noqa: E402,F401
from mau_b200 import data as D  # noqa: E402
from oracle import dataset_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=64)
    ap.add_argument("--edge", type=int, default=250)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--threads", default="1,2,4,8,0")
    ap.add_argument("--dir", default=None)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--no-crc", action="store_true")
    ap.add_argument("--zlib", action="store_true", help="inflate through zlib instead of the reader's own decoder")
    ap.add_argument("--stored", action="store_true", help="archives written with np.savez (stored members, no deflate)")
    ap.add_argument("--stage-bf16", action="store_true", help="also convert every decoded tile to the engine's staged bf16 NHWC layout in the decode pass")
    a = ap.parse_args()

    root = a.dir or tempfile.mkdtemp(prefix="mau_tiles_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    made = not (a.dir and os.path.isdir(os.path.join(a.dir, "train")))
    if made:
        t0 = time.perf_counter()
        O.write_synthetic_split(root, "train", a.tiles, a.edge, a.edge, seed=11, compressed=not a.stored)
        t_write = time.perf_counter() - t0
    files = O.list_split(root, "train")
    archive = sum(os.path.getsize(f) for f in files)
    dev = torch.device(a.device)

    for f in files:     # page cache warm for both paths
        with open(f, "rb") as fh:
            fh.read()
    O.collate([O.load_sample(f) for f in files[:2]])
    # reference path: one thread, np.load per sample (src/dataset.py:43-73), collate (src/dataset.py:87-108)
    t0 = time.perf_counter()
    n_ref = 0
    for b in range(0, len(files), a.batch):
        batch = O.collate([O.load_sample(f) for f in files[b:b + a.batch]], device=dev)
        n_ref += batch[0].shape[0]
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t_ref = time.perf_counter() - t0

    rows = []
    for th in [int(x) for x in a.threads.split(",")]:
        ds = D.FuturePredictionDataset("train", processed_dir=root, threads=th, verify_crc=not a.no_crc, use_zlib=a.zlib)
        loader = D.TileLoader(ds, a.batch, False, device=dev, prefetch=2, stage_bf16=a.stage_bf16)
        for _ in loader:        # warm-up epoch: page cache, staging ring
            pass
        s0 = ds.stats()
        t0 = time.perf_counter()
        n = 0
        for _ in range(a.epochs):
            for batch in loader:
                n += batch[0].shape[0]
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        s1 = ds.stats()
        rows.append({"threads": ds.threads, "tiles_per_s": round(n / dt, 1),
                     "payload_GBps": round((s1["payload_bytes"] - s0["payload_bytes"]) / dt / 1e9, 3),
                     "archive_MBps": round((s1["archive_bytes"] - s0["archive_bytes"]) / dt / 1e6, 1)})
        ds.close()
    out = {"bench": "tile loader", "tile": [23, a.edge, a.edge], "tiles": len(files), "batch": a.batch, "device": str(dev),
           "archive_bytes_per_tile": archive // len(files), "payload_bytes_per_tile": (25 * a.edge * a.edge + 4 + 828) * 4,
           "host_cores": os.cpu_count(), "crc": not a.no_crc, "members": "stored" if a.stored else "deflate", "inflate": "zlib" if a.zlib else "own decoder", "stage_bf16": a.stage_bf16,
           "reference_path": {"tiles_per_s": round(n_ref / t_ref, 1), "threads": 1, "what": "oracle/dataset_oracle.py: np.load + stack + pad_sequence"},
           "native": rows}
    print(json.dumps(out))
    if made and not a.dir:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
