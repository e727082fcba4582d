"""CPU oracle for the input pipeline: NumPy / torch restatement of the reference's
``src/dataset.py``.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; ``tests/`` use it as the
checker and ``tools/loader_bench.py`` times it as the CPU baseline of the loader.

It restates, without importing the reference: the per-sample read (``np.load`` of the four members, dates
from the file name, optional transform, ``.float()``: src/dataset.py:43-73), the collate step (stack,
``pad_sequence(batch_first=True, padding_value=0.0)``: src/dataset.py:87-108) and ``RandomFlip``
(src/dataset.py:134-141).  The synthetic tile generator follows SURVEY 8d and writes archives exactly like
the reference's writer (``np.savez_compressed`` of float32 arrays, src/data/processing_10m/process.py:187).

Parity status: pinned.  ``oracle/gen_golden_tiles.py`` imports the *real* ``src/dataset.py`` in the build
container (with a stand-in for the Hydra ``CONFIG`` object, which needs packages that are absent) and writes
``tests/golden/tiles/``: six small archives plus the batches the reference's own ``DataLoader`` yields for
them (sequential, shuffled under ``torch.manual_seed``, with ``RandomFlip``).  ``tests/test_tiles_cpu.py``
checks this restatement and the native reader against those batches.
"""
from __future__ import annotations

import os
import random
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


def sample_name(city: str, city_id: int, lat: float, lon: float, t1: Sequence[int], t2: Sequence[int]) -> str:
    """File-name layout the reference parses from the end (src/dataset.py:45-52, 76-84):
    ``<city>_<id>_<lat>_<lon>_<t1 year>_<t1 month>_to_<t2 year>_<t2 month>.npz``."""
    return f"{city}_{city_id}_{lat}_{lon}_{t1[0]}_{t1[1]}_to_{t2[0]}_{t2[1]}.npz"


def load_sample(filepath: str, transform: Optional[Callable] = None):
    """``FuturePredictionDataset.__getitem__`` (src/dataset.py:43-73)."""
    parts = os.path.basename(filepath).split("_")
    t1 = [int(parts[-5]), int(parts[-4])]
    t2 = [int(parts[-2]), int(parts[-1].split(".")[0])]
    with np.load(filepath) as data:
        x, y, md, ts = data["input"], data["target"], data["metadata"], data["temperature_serie"]
    if transform is not None:
        x, y = transform(x, y)
    return (torch.from_numpy(x).float(), torch.from_numpy(md).float(), torch.from_numpy(ts).float(),
            torch.tensor(t1).float(), torch.tensor(t2).float(), torch.from_numpy(y).float())


def collate(batch, device="cpu"):
    """``collate_fn`` (src/dataset.py:87-108)."""
    batch = [b for b in batch if b[0] is not None]
    if not batch:
        return tuple(torch.tensor([]) for _ in range(7))
    xs, mds, tss, t1s, t2s, ys = zip(*batch)
    lengths = torch.tensor([len(t) for t in tss])
    padded = torch.nn.utils.rnn.pad_sequence(list(tss), batch_first=True, padding_value=0.0).float()
    st = lambda s: torch.stack(s).float().to(device)  # noqa: E731
    return st(xs), st(mds), padded.to(device), lengths, st(t1s), st(t2s), st(ys)


class RandomFlip:
    """src/dataset.py:134-141."""

    def __init__(self, seed: int = 42):
        random.seed(seed)

    def __call__(self, x, y):
        if random.random() < 0.5:
            x = np.flip(x, axis=2).copy()
            y = np.flip(y, axis=2).copy()
        return x, y


def list_split(processed_dir: str, split: str) -> List[str]:
    """File list of one split, sorted (src/dataset.py:35-36)."""
    d = os.path.join(processed_dir, split)
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".npz"))


def synthetic_tile(rng: np.random.Generator, h: int, w: int, t: int = 828, channels: int = 23):
    """One sample of the SURVEY 8d shape: 9 one-hot Dynamic World planes, 3 normalised RGB, raw NDVI, normalised
    LST, 9 one-hot planes of a second class map that differs in ~10 % of the pixels; target = (NDVI, LST)."""
    cls1 = rng.integers(0, 9, (h, w))
    cls2 = np.where(rng.random((h, w)) < 0.1, rng.integers(0, 9, (h, w)), cls1)
    x = np.zeros((channels, h, w), np.float32)
    x[:9] = (cls1[None] == np.arange(9)[:, None, None])
    x[9:12] = rng.standard_normal((3, h, w))
    x[12] = rng.uniform(-1, 1, (h, w))
    x[13] = rng.standard_normal((h, w))
    x[14:23] = (cls2[None] == np.arange(9)[:, None, None])
    y = np.stack([rng.uniform(-1, 1, (h, w)), rng.standard_normal((h, w))]).astype(np.float32)
    md = rng.standard_normal(4).astype(np.float32)
    ts = rng.standard_normal(t).astype(np.float32)
    return x, y, md, ts


def write_synthetic_split(processed_dir: str, split: str, n: int, h: int, w: int, seed: int = 0, t_range=(790, 828),
                          compressed: bool = True) -> List[str]:
    """n archives like the reference's writer makes them (src/data/processing_10m/process.py:187)."""
    d = os.path.join(processed_dir, split)
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        x, y, md, ts = synthetic_tile(rng, h, w, int(rng.integers(t_range[0], t_range[1] + 1)))
        name = sample_name(f"Synth City{i % 7}", 1000 + i, round(float(rng.uniform(-60, 60)), 4), round(float(rng.uniform(-180, 180)), 4),
                           (2016 + i % 4, 1 + i % 12), (2020 + i % 4, 1 + (i * 5) % 12))
        p = os.path.join(d, name)
        (np.savez_compressed if compressed else np.savez)(p, input=x, target=y, metadata=md, temperature_serie=ts)
        out.append(p)
    return sorted(out)
