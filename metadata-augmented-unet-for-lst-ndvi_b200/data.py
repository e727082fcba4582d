"""Input pipeline in front of the hot path: host-side mirror of reference
``src/dataset.py`` on top of the native tile reader ``libmau_tiles.so`` (include/mau_tiles.h).

Same names and meaning as the reference:

* :class:`FuturePredictionDataset` (src/dataset.py:18-84) -- the sorted ``*.npz`` files of one split;
  ``__getitem__`` returns ``(input, metadata, temp_series, t1_date, t2_date, target)`` fp32 tensors,
  dates parsed from the file name.
* :func:`collate_fn` (src/dataset.py:87-108) -- stack, zero-pad the series, move to the device; returns
  ``(inputs, metadatas, temp_series_padded, temp_series_lengths, t1_dates, t2_dates, targets)``.
* :func:`create_dataloader` (src/dataset.py:110-131) -- iterable over such batches.
* :class:`RandomFlip` (src/dataset.py:134-141).

What differs is how a batch is produced.  The reference inflates four ZIP members per sample with
``np.load`` on the training thread (``num_workers=0`` because its collate function moves tensors to the
device, src/train.py:180).  Here the batch sampler's indices go to the native reader, which decodes the
samples of a batch in parallel on a pool of host threads straight into pinned batch buffers (stack, pad
and flip happen in that same pass); batches are prefetched ``prefetch`` deep and the host->device copies
run on a side stream, so that decoding, PCIe and the GPU step overlap.  (A variant with a dedicated staging
thread that issues each copy the moment its batch is decoded measured slower on the 16-core B200 box -- 842 vs
1 210 training tiles/s from disk -- and was dropped.)  Batch order, flip decisions and
values are identical to the reference loader under the same seeds (tests/test_tiles_cpu.py).

There is no NumPy fallback: a missing ``libmau_tiles.so`` raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import random
import threading
from collections import deque
from typing import Callable, List, Optional, Sequence

import torch
from torch.utils.data import BatchSampler, Dataset, RandomSampler, SequentialSampler

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmau_tiles.so")

E_ARG, E_IO, E_FORMAT, E_MEMBER, E_SHAPE, E_CAPACITY, E_DTYPE = 1, 2, 3, 4, 5, 6, 7
FLAG_NO_CRC, FLAG_ZLIB, FLAG_NICE = 1, 2, 4

# every symbol include/mau_tiles.h declares (tests check the .so exports all of them)
EXPORTS = ("mau_tiles_last_error", "mau_tiles_version", "mau_tiles_open", "mau_tiles_close", "mau_tiles_count",
           "mau_tiles_threads", "mau_tiles_probe", "mau_tiles_series_lengths", "mau_tiles_read_batch", "mau_tiles_submit", "mau_tiles_submit_staged",
           "mau_tiles_wait",
           "mau_tiles_done", "mau_tiles_repack", "mau_tiles_inflate", "mau_tiles_crc32", "mau_tiles_stats")

_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load libmau_tiles.so once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"mau_b200: tile reader not built ({LIB_PATH} missing). Run "
                               "`python -c 'import __graft_entry__ as g; g.build()'` -- there is no NumPy fallback.")
        L = C.CDLL(LIB_PATH)
        p64, pf, pu8 = C.POINTER(C.c_int64), C.c_void_p, C.c_void_p
        L.mau_tiles_last_error.restype = C.c_char_p
        L.mau_tiles_version.restype = C.c_int
        L.mau_tiles_open.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.mau_tiles_close.argtypes = [C.c_void_p]
        L.mau_tiles_count.argtypes = [C.c_void_p]
        L.mau_tiles_count.restype = C.c_int64
        L.mau_tiles_threads.argtypes = [C.c_void_p]
        L.mau_tiles_probe.argtypes = [C.c_void_p, C.c_int64, p64]
        L.mau_tiles_series_lengths.argtypes = [C.c_void_p, p64, C.c_int64, p64]
        batch = [C.c_void_p, p64, C.c_int64, pu8, p64, pf, pf, pf, pf, C.c_int64, C.c_void_p]
        L.mau_tiles_read_batch.argtypes = batch
        L.mau_tiles_submit.argtypes = batch
        L.mau_tiles_submit.restype = C.c_int64
        L.mau_tiles_submit_staged.argtypes = [C.c_void_p, p64, C.c_int64, pu8, p64, pf, C.c_void_p, C.c_int64, pf, pf, pf, C.c_int64, C.c_void_p]
        L.mau_tiles_submit_staged.restype = C.c_int64
        L.mau_tiles_wait.argtypes = [C.c_void_p, C.c_int64]
        L.mau_tiles_done.argtypes = [C.c_void_p, C.c_int64]
        L.mau_tiles_stats.argtypes = [C.c_void_p, p64, p64, p64]
        L.mau_tiles_inflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]
        L.mau_tiles_repack.argtypes = [C.c_char_p, C.c_char_p]
        L.mau_tiles_crc32.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        L.mau_tiles_crc32.restype = C.c_uint32
        _lib = L
    return _lib


def _raise(code: int):
    """Map a MAU_TILES_E_* code onto the exception type the reference's np.load / torch.stack path raises."""
    msg = (lib().mau_tiles_last_error() or b"").decode("utf-8", "replace")
    if code == E_MEMBER:
        raise KeyError(msg)                       # NpzFile.__getitem__
    if code == E_IO:
        raise FileNotFoundError(msg) if "No such file" in msg else OSError(msg)
    if code in (E_SHAPE, E_CAPACITY):
        raise RuntimeError(msg)                   # torch.stack on unequal shapes
    if code == E_ARG:
        raise IndexError(msg) if "out of range" in msg else ValueError(msg)
    raise ValueError(msg)                         # BadZipFile / unpicklable member / dtype


def repack_split(src_dir: str, dst_dir: str, workers: int = 0) -> int:
    """Rewrites every ``*.npz`` of ``src_dir`` into ``dst_dir`` with stored (uncompressed) members -- the format
    ``np.savez`` writes, still readable by ``np.load`` and by the reference's loader.  Decoding such archives is a
    memcpy instead of an inflate: on hosts where the decode, not the GPU, bounds training from disk,
    trade 3.5x the disk space for it once.  Returns the number of archives written."""
    from concurrent.futures import ThreadPoolExecutor
    names = sorted(f for f in os.listdir(src_dir) if f.endswith(".npz"))
    os.makedirs(dst_dir, exist_ok=True)
    L = lib()

    def one(name):
        rc = L.mau_tiles_repack(os.fsencode(os.path.join(src_dir, name)), os.fsencode(os.path.join(dst_dir, name)))
        if rc:
            _raise(rc)      # last_error is thread-local and this is the thread that made the call

    with ThreadPoolExecutor(max_workers=workers or os.cpu_count() or 4) as ex:
        list(ex.map(one, names))
    return len(names)


def parse_dates(filename: str):
    """``<city words>_<id>_<lat>_<lon>_<t1 year>_<t1 month>_<..>_<t2 year>_<t2 month>.npz`` ->
    (t1_year, t1_month, t2_year, t2_month), the fields src/dataset.py:47-52 reads (5th, 4th, 2nd and last
    from the end; the last one up to its first '.').  Raises like ``int()`` / list indexing do."""
    parts = os.path.basename(filename).split("_")
    return int(parts[-5]), int(parts[-4]), int(parts[-2]), int(parts[-1].split(".")[0])


def _resolve_processed_dir(processed_dir: Optional[str]) -> str:
    if processed_dir is not None:
        return str(processed_dir)
    env = os.environ.get("MAU_PROCESSED_IMAGE_DATASET")
    if env:
        return env
    try:    # the reference's global config (needs hydra/omegaconf), src/dataset.py:27
        from urban_planner.config import CONFIG  # type: ignore
        return str(CONFIG.PROCESSED_IMAGE_DATASET)
    except Exception as e:
        raise RuntimeError("FuturePredictionDataset: pass processed_dir=..., set MAU_PROCESSED_IMAGE_DATASET, or make "
                           f"urban_planner.config importable ({type(e).__name__}: {e})")


def _resolve_device(device) -> torch.device:
    if device is not None:
        return torch.device(device)
    try:    # src/dataset.py:88
        from urban_planner.config import CONFIG  # type: ignore
        return torch.device(CONFIG.device)
    except Exception:
        return torch.device("cuda:0" if torch.cuda.is_available() else "cpu")


class RandomFlip:
    """Horizontal flip of input and target with probability 0.5 (src/dataset.py:134-141).  Python's global
    ``random`` is seeded at construction and consulted once per sample, exactly like the reference, so the
    decisions are the same sequence.  The loader recognises this class and performs the flip inside the native
    decode pass (:meth:`decide`); called directly it flips NumPy arrays like the reference's transform."""

    def __init__(self, seed: Optional[int] = None):
        if seed is None:
            try:
                from urban_planner.config import CONFIG  # type: ignore
                seed = int(CONFIG.seed)
            except Exception:
                seed = 42          # conf/config.yaml:61
        random.seed(seed)

    @staticmethod
    def decide() -> bool:
        return random.random() < 0.5

    def __call__(self, x, y):
        if self.decide():
            import numpy as np
            x = np.flip(x, axis=2).copy()
            y = np.flip(y, axis=2).copy()
        return x, y


class FuturePredictionDataset(Dataset):
    """The ``*.npz`` samples of one split (src/dataset.py:18-84), read through the native tile reader.

    ``processed_dir`` replaces ``CONFIG.PROCESSED_IMAGE_DATASET`` when the reference's Hydra config is not
    importable; ``threads`` sizes the decode pool (default: all online cores); ``verify_crc=False`` skips the
    CRC-32 check ``zipfile`` performs; ``use_zlib=True`` inflates through zlib instead of the reader's own
    DEFLATE decoder (A/B switch for benchmarks)."""

    SERIES_CAPACITY = 1024      # >= conf/config.yaml:20 seq_len 828; grown on demand

    def __init__(self, split: str, transform: Optional[Callable] = None, processed_dir: Optional[str] = None,
                 threads: int = 0, verify_crc: bool = True, use_zlib: bool = False):
        self.processed_dir = _resolve_processed_dir(processed_dir)
        self.split = split
        self.transform = transform
        self.data_dir = os.path.join(self.processed_dir, self.split)
        if not os.path.isdir(self.data_dir):
            raise FileNotFoundError(f"Directory for split '{self.split}' not found at: {self.data_dir}")
        self.file_list = sorted(os.path.join(self.data_dir, f) for f in os.listdir(self.data_dir) if f.endswith(".npz"))
        self._threads, self._flags = int(threads), (0 if verify_crc else FLAG_NO_CRC) | (FLAG_ZLIB if use_zlib else 0)
        self._h, self._h_pid = None, None
        self._dims: Optional[List[int]] = None
        self._series_capacity = self.SERIES_CAPACITY
        lib()                       # fail at construction if the reader is not built

    # -- plumbing ------------------------------------------------------------------------------------------
    @property
    def _handle(self):
        """The native reader (file list + worker pool) of *this process*.  Opened on first use and re-opened after a
        fork: threads do not survive ``fork()``, so a dataset handed to ``torch.utils.data.DataLoader(num_workers>0)``
        gets a fresh pool in every worker instead of waiting on one that no longer exists."""
        if self._h is None or self._h_pid != os.getpid():
            h = C.c_void_p()
            arr = (C.c_char_p * len(self.file_list))(*[os.fsencode(p) for p in self.file_list])
            rc = lib().mau_tiles_open(arr, len(self.file_list), self._threads, self._flags, C.byref(h))
            if rc:
                _raise(rc)
            self._h, self._h_pid = h, os.getpid()     # a handle inherited from the parent is abandoned, not closed:
        return self._h                                # closing it would join threads that do not exist here

    def __getstate__(self):          # spawn / forkserver workers, copy.deepcopy: the handle stays with its process
        d = dict(self.__dict__)
        d["_h"], d["_h_pid"] = None, None
        return d

    def __del__(self):
        h, pid = getattr(self, "_h", None), getattr(self, "_h_pid", None)
        self._h = self._h_pid = None
        if h is not None and pid == os.getpid():
            try:
                lib().mau_tiles_close(h)
            except Exception:
                pass

    def close(self):
        self.__del__()

    @property
    def threads(self) -> int:
        return int(lib().mau_tiles_threads(self._handle))

    def probe(self, idx: int) -> List[int]:
        """[input C, H, W, target C, H, W, metadata length, series length] from the NPY headers of sample idx."""
        d = (C.c_int64 * 8)()
        rc = lib().mau_tiles_probe(self._handle, int(idx), d)
        if rc:
            _raise(rc)
        return list(d)

    def series_lengths(self, indices: Sequence[int]) -> List[int]:
        """Series lengths of ``indices`` from the NPY headers (decoded on the pool, cached: they never change)."""
        cache = self.__dict__.setdefault("_series_len", {})
        missing = [int(i) for i in dict.fromkeys(indices) if i not in cache]
        if missing:
            n = len(missing)
            idx, out = (C.c_int64 * n)(*missing), (C.c_int64 * n)()
            rc = lib().mau_tiles_series_lengths(self._handle, idx, n, out)
            if rc:
                _raise(rc)
            cache.update(zip(missing, out))
        return [cache[int(i)] for i in indices]

    def batch_dims(self) -> List[int]:
        """Shapes every batch is checked against: those of sample 0 (all samples of a processed dataset share them,
        src/data/processing_10m/process.py:165-187)."""
        if self._dims is None:
            self._dims = self.probe(0)
        return self._dims

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        lib().mau_tiles_stats(self._handle, C.byref(a), C.byref(b), C.byref(c))
        return {"payload_bytes": a.value, "archive_bytes": b.value, "samples": c.value}

    # -- staging buffers -----------------------------------------------------------------------------------
    def alloc_staging(self, n: int, pin: bool = False, dims: Optional[Sequence[int]] = None, stage_bf16: bool = False):
        d = list(dims) if dims is not None else self.batch_dims()
        kw = dict(dtype=torch.float32, pin_memory=bool(pin))
        st = self._alloc_staging(n, d, kw)
        if stage_bf16:      # the input tiles a second time in the engine's staged layout (bf16 NHWC, channel stride padded to 8)
            st["input_staged"] = torch.empty((n, d[1], d[2], (d[0] + 7) // 8 * 8), dtype=torch.bfloat16, pin_memory=bool(pin))
        return st

    def _alloc_staging(self, n, d, kw):
        return {
            "dims": d,
            "input": torch.empty((n, d[0], d[1], d[2]), **kw),
            "target": torch.empty((n, d[3], d[4], d[5]), **kw),
            "metadata": torch.empty((n, d[6]), **kw),
            "series": torch.empty((n, self._series_capacity), **kw),
            "series_len": torch.empty((n,), dtype=torch.int64),
            "dates": torch.empty((n, 4), **kw),
            "capacity": n,
        }

    def submit(self, indices: Sequence[int], flips: Optional[Sequence[bool]], st, min_width: int = 0) -> "_Ticket":
        """Start decoding ``indices`` into staging set ``st`` (non-blocking); dates are parsed here, on the caller's
        thread, so that malformed file names raise where the reference raises them (``__getitem__``)."""
        n = len(indices)
        if n > st["capacity"]:
            raise ValueError(f"batch of {n} exceeds the staging capacity {st['capacity']}")
        if n:
            st["dates"][:n] = torch.tensor([parse_dates(self.file_list[i]) for i in indices], dtype=torch.float32)
        idx = (C.c_int64 * n)(*[int(i) for i in indices])
        fl = (C.c_uint8 * n)(*[1 if f else 0 for f in flips]) if flips is not None else None
        dims = (C.c_int64 * 8)(*st["dims"])
        staged = st.get("input_staged")
        if staged is not None and self.transform is not None and not isinstance(self.transform, RandomFlip):
            raise ValueError("stage_bf16 converts the tiles inside the decode pass: only RandomFlip (applied there too) is "
                             "supported as a transform, an arbitrary callable would run on the fp32 copy only")
        ticket = lib().mau_tiles_submit_staged(self._handle, idx, n, fl, dims, st["input"].data_ptr(),
                                               staged.data_ptr() if staged is not None else None,
                                               staged.shape[3] if staged is not None else 0, st["target"].data_ptr(),
                                               st["metadata"].data_ptr(), st["series"].data_ptr(), st["series"].shape[1],
                                               st["series_len"].data_ptr())
        if ticket < 0:
            _raise(int(-ticket))
        return _Ticket(self, ticket, list(indices), list(flips) if flips is not None else None, st, n, min_width)

    def read_batch(self, indices: Sequence[int], flips: Optional[Sequence[bool]] = None, st=None):
        """Blocking decode of one batch; returns the staging set (views valid until it is reused)."""
        st = st if st is not None else self.alloc_staging(max(len(indices), 1))
        return self.submit(indices, flips, st).wait()

    # -- reference Dataset protocol --------------------------------------------------------------------------
    def __len__(self):
        return len(self.file_list)

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError("list index out of range")
        y1, m1, y2, m2 = parse_dates(self.file_list[idx])        # before the read, like src/dataset.py:47-54
        native_flip = isinstance(self.transform, RandomFlip)
        st = self.read_batch([idx], None, self.alloc_staging(1, dims=self.probe(idx)))
        n_series = int(st["series_len"][0])
        x, y = st["input"][0], st["target"][0]      # an arbitrary transform has already run on the staging set (_Ticket.wait)
        if native_flip and self.transform.decide():
            x, y = x.flip(2), y.flip(2)
        return (x.contiguous(), st["metadata"][0].clone(), st["series"][0, :n_series].clone(),
                torch.tensor([y1, m1]).float(), torch.tensor([y2, m2]).float(), y.contiguous())

    def get_metadata_from_idx(self, idx: int) -> dict:
        """City, latitude and longitude from the file name (src/dataset.py:76-84)."""
        parts = os.path.basename(self.file_list[idx]).split("_")
        return {"city": " ".join(parts[:-8]), "lat": float(parts[-7]), "lon": float(parts[-6])}


class _Ticket:
    def __init__(self, ds: FuturePredictionDataset, ticket: int, indices, flips, st, n: int, min_width: int = 0):
        self.ds, self.ticket, self.indices, self.flips, self.st, self.n = ds, ticket, indices, flips, st, n
        self.min_width = min_width

    def done(self) -> bool:
        return lib().mau_tiles_done(self.ds._handle, self.ticket) != 0

    def wait(self):
        rc = lib().mau_tiles_wait(self.ds._handle, self.ticket)
        if rc == E_CAPACITY:        # a series longer than the staging row: grow and decode this batch again
            ds, st = self.ds, self.st
            longest = max(ds.series_lengths(self.indices))
            ds._series_capacity = max(ds._series_capacity, 2 * longest)
            st["series"] = torch.empty((st["capacity"], ds._series_capacity), dtype=torch.float32,
                                       pin_memory=st["series"].is_pinned())
            return ds.submit(self.indices, self.flips, st, self.min_width).wait()
        if rc:
            _raise(rc)
        self.st["n"], self.st["min_width"] = self.n, self.min_width
        tf = self.ds.transform
        if tf is not None and not isinstance(tf, RandomFlip):
            _apply_generic_transform(tf, self.st)
        return self.st


def _apply_generic_transform(tf: Callable, st):
    """An arbitrary ``transform(input, target) -> (input, target)`` on NumPy arrays (src/dataset.py:61-62), applied
    per sample in batch order on the consumer's thread; shapes must be preserved (torch.stack would fail otherwise)."""
    for k in range(st["n"]):
        xa, ya = tf(st["input"][k].numpy(), st["target"][k].numpy())
        xt, yt = torch.as_tensor(xa.copy()).float(), torch.as_tensor(ya.copy()).float()
        if xt.shape != st["input"][k].shape or yt.shape != st["target"][k].shape:
            raise RuntimeError(f"stack expects each tensor to be equal size, but the transform returned "
                               f"{list(xt.shape)} / {list(yt.shape)}")
        st["input"][k].copy_(xt)
        st["target"][k].copy_(yt)


def _batch_from_staging(st, device: torch.device, non_blocking: bool = False):
    """The 7-tuple of collate_fn (src/dataset.py:108) from a decoded staging set."""
    n = st["n"]
    lengths = st["series_len"][:n].clone()
    width = max(int(lengths.max()) if n else 0, int(st.get("min_width", 0)))       # min_width: longest series of the global batch
    if width > st["series"].shape[1]:
        raise RuntimeError(f"series width {width} exceeds the staging row {st['series'].shape[1]}")
    to = dict(device=device, non_blocking=non_blocking)
    # stage_bf16: ship the bf16 NHWC tiles (48 instead of 92 bytes per pixel); the model accepts them as `maps`
    inputs = (st["input_staged"] if "input_staged" in st else st["input"])[:n].to(**to)
    metadatas = st["metadata"][:n].to(**to)
    targets = st["target"][:n].to(**to)
    dates = st["dates"][:n].to(**to)
    series = st["series"][:n].to(**to)[:, :width]
    if device.type == "cpu":       # .to() on the same device aliases the staging set, which the loader reuses
        inputs, metadatas, targets, dates, series = inputs.clone(), metadatas.clone(), targets.clone(), dates.clone(), series.clone()
    series = series.contiguous()
    return inputs, metadatas, series, lengths, dates[:, 0:2].contiguous(), dates[:, 2:4].contiguous(), targets


def collate_fn(batch, device=None):
    """Reference collate (src/dataset.py:87-108) for a list of ``__getitem__`` tuples: drops samples whose input is
    None, stacks, zero-pads the series to the longest of the batch and moves everything to the device.  Used with
    plain ``torch.utils.data.DataLoader``; :func:`create_dataloader` never materialises per-sample tuples."""
    device = _resolve_device(device)
    batch = [b for b in batch if b[0] is not None]
    if not batch:
        return tuple(torch.tensor([]) for _ in range(7))
    inputs, metadatas, temp_series, t1_dates, t2_dates, targets = zip(*batch)
    lengths = torch.tensor([len(ts) for ts in temp_series])
    width = int(lengths.max())
    padded = torch.zeros((len(batch), width) + tuple(temp_series[0].shape[1:]), dtype=torch.float32)
    for k, ts in enumerate(temp_series):
        padded[k, :len(ts)] = ts
    mv = lambda seq: torch.stack(seq).float().to(device)  # noqa: E731
    return mv(inputs), mv(metadatas), padded.to(device), lengths, mv(t1_dates), mv(t2_dates), mv(targets)


class TileLoader:
    """Iterable over collated device batches -- what ``create_dataloader`` returns in place of the reference's
    ``DataLoader(dataset, batch_size, shuffle, num_workers=0, collate_fn=collate_fn)`` (src/dataset.py:124-130).

    Index order comes from torch's own ``RandomSampler`` / ``SequentialSampler`` + ``BatchSampler`` (the classes
    ``DataLoader`` instantiates), so a given ``torch.manual_seed`` yields the reference's batches.  With
    ``world_size > 1`` every rank draws the same global batch of ``batch_size * world_size`` indices and keeps its
    contiguous slice (all ranks must share the torch seed, as src/train.py:109 sets it)."""

    def __init__(self, dataset: FuturePredictionDataset, batch_size: int, shuffle: bool, device=None, prefetch: int = 2,
                 drop_last: bool = False, rank: int = 0, world_size: int = 1, generator=None, stage_bf16: bool = False):
        if batch_size <= 0 or world_size <= 0 or not 0 <= rank < world_size:
            raise ValueError("batch_size and world_size must be positive, rank in [0, world_size)")
        self.dataset, self.batch_size, self.shuffle = dataset, int(batch_size), bool(shuffle)
        self.device = _resolve_device(device)
        self.prefetch = max(1, int(prefetch))
        self.stage_bf16 = bool(stage_bf16)      # yield `inputs` as the engine's staged bf16 NHWC tiles (halves the PCIe bytes)
        self.rank, self.world_size = int(rank), int(world_size)
        self.sampler = RandomSampler(dataset, generator=generator) if shuffle else SequentialSampler(dataset)
        self.batch_sampler = BatchSampler(self.sampler, self.batch_size * self.world_size, drop_last)
        self.generator = generator
        self._pin = self.device.type == "cuda"
        self._rings = None

    def __len__(self):
        return len(self.batch_sampler)

    def _local(self, global_batch: List[int]) -> List[int]:
        if self.world_size == 1:
            return global_batch
        per = -(-len(global_batch) // self.world_size)        # the last global batch may be short
        return global_batch[self.rank * per:(self.rank + 1) * per]

    def _flips(self, global_batch: List[int]) -> Optional[List[bool]]:
        """One ``random.random()`` per sample of the *global* batch in sample order (every rank replays the same
        sequence and keeps its slice), which is the order the reference's ``__getitem__`` calls consume them in."""
        tf = self.dataset.transform
        if tf is None:
            return None
        if isinstance(tf, RandomFlip):
            return self._local([tf.decide() for _ in global_batch])
        return None        # any other callable runs per sample after the decode (_apply_generic_transform)

    def __iter__(self):
        ds = self.dataset
        # prefetch + 2 staging sets: up to two being copied H2D, `prefetch` being decoded.  They are kept between epochs
        # (pinning 100 MB buffers is slow) but owned by one iterator at a time: a second, concurrent iterator over the
        # same loader finds none cached and allocates its own
        if len(self.batch_sampler) == 0:      # an empty split yields nothing (and has no sample to take shapes from)
            return
        rings, self._rings = self._rings, None
        if rings is None or any(st["series"].shape[1] < ds._series_capacity for st in rings):
            rings = [ds.alloc_staging(self.batch_size, pin=self._pin, stage_bf16=self.stage_bf16) for _ in range(self.prefetch + 2)]
        free = deque(rings)
        busy = {}                     # id(staging set) -> event of its last H2D copy
        inflight = deque()            # decode tickets, oldest first
        staged = deque()              # (device batch, copy event) ahead of the consumer, at most one
        copy_stream = torch.cuda.Stream(device=self.device) if self._pin else None
        # DataLoader's iterator draws its base seed from the default generator before the sampler draws the
        # permutation seed (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__); replaying that draw keeps
        # the batch order identical to the reference loader under the same torch.manual_seed
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)
        it = iter(self.batch_sampler)

        def top_up():
            while free and len(inflight) < self.prefetch:
                gb = next(it, None)
                if gb is None:
                    return
                flips = self._flips(gb)
                local = self._local(gb)
                # the reference pads every series to the longest of the batch and the LSTM runs over that padding
                # (src/dataset.py:106, src/model.py:29-33): ranks pad to the longest series of the *global* batch
                width = max(ds.series_lengths(gb)) if self.world_size > 1 else 0
                if width > ds._series_capacity:
                    ds._series_capacity = 2 * width       # staging sets of the next epoch are allocated at this capacity
                st = free.popleft()
                ev = busy.pop(id(st), None)
                if ev is not None:
                    ev.synchronize()          # the copy that last read this staging set has finished
                if st["series"].shape[1] < width:
                    st["series"] = torch.empty((st["capacity"], ds._series_capacity), dtype=torch.float32, pin_memory=self._pin)
                inflight.append(ds.submit(local, flips, st, width))

        def stage(block: bool) -> bool:
            """Move the oldest decoded batch to the device (asynchronously on the copy stream)."""
            if not inflight or (not block and not inflight[0].done()):
                return False
            st = inflight.popleft().wait()
            if copy_stream is not None:
                with torch.cuda.stream(copy_stream):
                    out = _batch_from_staging(st, self.device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                busy[id(st)] = ev
            else:
                out, ev = _batch_from_staging(st, self.device), None
            staged.append((out, ev))
            free.append(st)
            top_up()
            return True

        try:
            top_up()
            while staged or stage(True):
                out, ev = staged.popleft()
                stage(False)          # the next batch's H2D starts now if it is decoded: it overlaps this batch's step
                if ev is not None:    # even when the caller synchronises every step (loss.item(), src/train.py:258)
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(ev)
                    for t in out:
                        if t.is_cuda:
                            t.record_stream(cur)
                yield out
        finally:
            while inflight:           # never leave the pool writing into buffers we are about to drop
                try:
                    inflight.popleft().wait()
                except Exception:
                    pass
            try:
                for ev in busy.values():      # copies still reading the staging sets we hand back
                    ev.synchronize()
                self._rings = rings
            except Exception:                 # e.g. a torn-down CUDA context at interpreter exit: just drop them
                pass


def create_dataloader(split: str, batch_size: int, shuffle: bool, dataset_type: str, transform=None, num_workers: int = 0,
                      *, device=None, processed_dir: Optional[str] = None, prefetch: int = 2, drop_last: bool = False,
                      rank: int = 0, world_size: int = 1, threads: Optional[int] = None, generator=None,
                      stage_bf16: bool = False) -> TileLoader:
    """src/dataset.py:110-131 with the same positional arguments.  ``num_workers`` sizes the native decode pool
    (0, the reference's value, means all online cores -- there are no worker *processes*)."""
    assert dataset_type == "future", "Only 'future' dataset_type is supported in create_dataloader."
    ds = FuturePredictionDataset(split=split, transform=transform, processed_dir=processed_dir,
                                 threads=int(threads if threads is not None else num_workers))
    return TileLoader(ds, batch_size, shuffle, device=device, prefetch=prefetch, drop_last=drop_last, rank=rank,
                      world_size=world_size, generator=generator, stage_bf16=stage_bf16)
