"""Writes tests/golden/tiles/ -- run in the build container, where /root/reference exists.

Six small sample archives in the reference's on-disk format (np.savez_compressed of float32 `input`, `target`,
`metadata`, `temperature_serie`; one archive is written with np.savez, i.e. stored members) and the batches the
REAL reference loader (`src/dataset.py`: FuturePredictionDataset + collate_fn + torch DataLoader + RandomFlip)
yields for them.  `urban_planner.config` needs hydra/omegaconf, which are absent here, so a stand-in module with
the three attributes src/dataset.py reads (PROCESSED_IMAGE_DATASET, device, seed) is installed first.

TEST INFRASTRUCTURE ONLY (see oracle/dataset_oracle.py)."""
import os
import shutil
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "tiles")
REF = "/root/reference"


def import_reference_dataset(processed_dir):
    cfg = types.SimpleNamespace(PROCESSED_IMAGE_DATASET=processed_dir, device="cpu", seed=42)
    pkg, mod = types.ModuleType("urban_planner"), types.ModuleType("urban_planner.config")
    mod.CONFIG = cfg
    pkg.config = mod
    sys.modules["urban_planner"], sys.modules["urban_planner.config"] = pkg, mod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.dataset as ref_dataset
    return ref_dataset


def write_archives():
    shutil.rmtree(OUT, ignore_errors=True)
    d = os.path.join(OUT, "train")
    os.makedirs(d)
    rng = np.random.default_rng(20261018)
    names = ["Ville Neuve_17_48.8566_2.3522_2017_6_to_2021_7.npz", "Aix en Provence_3_43.5297_5.4474_2016_12_to_2020_1.npz",
             "Lyon_250_45.764_4.8357_2018_1_to_2022_11.npz", "Porto_9_41.1579_-8.6291_2019_7_to_2023_7.npz",
             "Sao Paulo Sul_1200_-23.5505_-46.6333_2016_3_to_2024_2.npz", "Oslo_77_59.9139_10.7522_2019_10_to_2025_9.npz"]
    for i, n in enumerate(names):
        cls = rng.integers(0, 9, (6, 10))
        x = np.zeros((23, 6, 10), np.float32)
        x[:9] = cls[None] == np.arange(9)[:, None, None]
        x[9:14] = rng.standard_normal((5, 6, 10))
        x[14:] = np.roll(x[:9], 1, axis=2)
        y = rng.standard_normal((2, 6, 10)).astype(np.float32)
        md = rng.standard_normal(4).astype(np.float32)
        ts = rng.standard_normal(5 + (i * 3) % 5).astype(np.float32)
        save = np.savez if i == 2 else np.savez_compressed
        save(os.path.join(d, n), input=x, target=y, metadata=md, temperature_serie=ts)


def main():
    write_archives()
    R = import_reference_dataset(OUT)
    keys = ("inputs", "metadatas", "series", "lengths", "t1", "t2", "targets")
    out = {}

    def record(tag, loader):
        nb = 0
        for b, batch in enumerate(loader):
            for k, t in zip(keys, batch):
                out[f"{tag}_b{b}_{k}"] = t.numpy()
            nb += 1
        out[f"{tag}_batches"] = np.array(nb)

    record("seq", R.create_dataloader("train", 4, False, "future"))
    torch.manual_seed(123)
    record("shuf", R.create_dataloader("train", 4, True, "future"))
    torch.manual_seed(7)
    loader = R.create_dataloader("train", 4, True, "future", transform=R.RandomFlip())
    record("flip_e0", loader)
    record("flip_e1", loader)          # second epoch: both RNG streams continue
    ds = R.FuturePredictionDataset("train")
    out["meta_from_idx"] = np.array([repr(ds.get_metadata_from_idx(i)) for i in range(len(ds))])
    np.savez_compressed(os.path.join(OUT, "expected_batches.npz"), **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
