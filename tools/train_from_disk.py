"""Training fed from .npz archives through the native tile loader on one GPU:

    python tools/train_from_disk.py [--tiles 96] [--batch 16] [--epochs 3] [--threads 0] [--stored]

Writes synthetic archives of the SURVEY 8d shape to /dev/shm, then runs BASELINE.json configs[2] (U-Net + metadata,
forward + L1 loss kernel + backward + fused AdamW) with every batch decoded from disk, stacked, copied H2D and
consumed -- i.e. the loop of src/train.py:235-258 including its per-step loss read-back.  Prints one JSON line:
tiles/s from disk, tiles/s of the loader alone (decode + H2D, no model) and of the model alone (resident inputs)."""
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

import mau_b200  # noqa: E402
from mau_b200 import data as D, engine  # noqa: E402
from oracle import dataset_oracle as O  # noqa: E402  (synthetic archive writer only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=96)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--prefetch", type=int, default=3)
    ap.add_argument("--stored", action="store_true")
    ap.add_argument("--zlib", action="store_true", help="inflate through zlib instead of the reader's own decoder")
    ap.add_argument("--edge", type=int, default=250)
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic tiles written (deflating them is slow); "
                    "the rest of --tiles are copies under other names -- the decode work per file is the same")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    root = tempfile.mkdtemp(prefix="mau_tiles_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        made = O.write_synthetic_split(root, "train", min(a.distinct, a.tiles), a.edge, a.edge, seed=11, compressed=not a.stored)
        for k in range(len(made), a.tiles):
            src = made[k % len(made)]
            head, tail = os.path.basename(src).split("_", 1)
            shutil.copyfile(src, os.path.join(os.path.dirname(src), f"{head} copy{k}_{tail}"))
        torch.manual_seed(42)
        model = mau_b200.UrbanPredictor("unet", 23, 828, 64, 8, 64, 96, 2, temporal_embeddings=False, metadata_embeddings=True).to(dev).train()
        opt = mau_b200.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-3)
        ds = D.FuturePredictionDataset("train", transform=D.RandomFlip(42), processed_dir=root, threads=a.threads, use_zlib=a.zlib)
        loader = D.TileLoader(ds, a.batch, True, device=dev, prefetch=a.prefetch, drop_last=True)

        def train_step(batch, read_back=True):
            inputs, metadatas, series, lengths, t1, t2, targets = batch
            md = torch.cat([metadatas, t1, t2], dim=1)                              # src/train.py:244
            out = model(inputs, series, md)
            loss = engine.compute_loss_l1_grad(out, targets, 0.0)["total"]
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss.item() if read_back else loss                                # src/train.py:258

        def epoch(fn):
            n = 0
            for batch in loader:
                fn(batch)
                n += batch[0].shape[0]
            torch.cuda.synchronize()
            return n

        keep = []
        epoch(lambda b: (train_step(b), keep.append(b) if len(keep) < 4 else None))        # warm-up: plans, allocator, page cache
        res = {}
        for name, fn in (("train_from_disk", train_step), ("loader_only", lambda b: None)):
            t0 = time.perf_counter()
            n = sum(epoch(fn) for _ in range(a.epochs))
            res[name] = n / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        steps = 3 * len(loader)
        for i in range(steps):
            train_step(keep[i % len(keep)])
        torch.cuda.synchronize()
        res["model_only_resident_inputs"] = steps * a.batch / (time.perf_counter() - t0)
        print(json.dumps({"bench": "training from .npz archives", "tile": [23, a.edge, a.edge], "batch": a.batch, "tiles": a.tiles,
                          "members": "stored" if a.stored else "deflate", "inflate": "zlib" if a.zlib else "own decoder", "decode_threads": loader.dataset.threads,
                          "host_cores": os.cpu_count(), "prefetch": a.prefetch,
                          "tiles_per_s": {k: round(v, 1) for k, v in res.items()}}))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
