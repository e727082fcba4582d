"""CPU: the native tile reader + loader (include/mau_tiles.h, mau_b200.data) against
(a) the batches the REAL reference loader produced for tests/golden/tiles (oracle/gen_golden_tiles.py),
(b) the NumPy restatement oracle/dataset_oracle.py on freshly written archives, bit-exact in both cases."""
import os
import struct
import zipfile

import numpy as np
import pytest
import torch

import mau_b200  # noqa: F401
from mau_b200 import data as D
from oracle import dataset_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiles")
KEYS = ("inputs", "metadatas", "series", "lengths", "t1", "t2", "targets")


@pytest.fixture(scope="module")
def expected():
    return np.load(os.path.join(GOLD, "expected_batches.npz"))


def assert_batches_equal(tag, expected, batches):
    assert len(batches) == int(expected[f"{tag}_batches"])
    for b, batch in enumerate(batches):
        for k, t in zip(KEYS, batch):
            ref = expected[f"{tag}_b{b}_{k}"]
            assert tuple(t.shape) == ref.shape, (tag, b, k, t.shape, ref.shape)
            assert str(t.dtype).replace("torch.", "") == str(ref.dtype), (tag, b, k, t.dtype, ref.dtype)
            assert np.array_equal(t.numpy(), ref), (tag, b, k)


# ---- the C ABI ---------------------------------------------------------------------------------------------
def test_library_exports_every_symbol_of_the_header():
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "..", "include", "mau_tiles.h")).read()
    import re
    declared = set(re.findall(r"\b(mau_tiles_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(D.EXPORTS), declared ^ set(D.EXPORTS)
    L = D.lib()
    assert all(hasattr(L, s) for s in D.EXPORTS)
    assert L.mau_tiles_version() >= 1


def test_probe_reports_npy_header_shapes():
    ds = D.FuturePredictionDataset("train", processed_dir=GOLD, threads=2)
    assert len(ds) == 6 and ds.threads == 2
    for i, f in enumerate(ds.file_list):
        with np.load(f) as z:
            want = list(z["input"].shape) + list(z["target"].shape) + [z["metadata"].shape[0], z["temperature_serie"].shape[0]]
        assert ds.probe(i) == want
    with pytest.raises(IndexError):
        ds.probe(6)


# ---- golden batches of the real reference loader -----------------------------------------------------------
def test_sequential_batches_equal_the_reference_loader(expected):
    loader = D.create_dataloader("train", 4, False, "future", device="cpu", processed_dir=GOLD)
    assert len(loader) == 2
    assert_batches_equal("seq", expected, list(loader))


def test_shuffled_batches_equal_the_reference_loader_under_the_same_torch_seed(expected):
    torch.manual_seed(123)
    loader = D.create_dataloader("train", 4, True, "future", device="cpu", processed_dir=GOLD)
    assert_batches_equal("shuf", expected, list(loader))


def test_random_flip_two_epochs_equal_the_reference_loader(expected):
    torch.manual_seed(7)
    loader = D.create_dataloader("train", 4, True, "future", transform=D.RandomFlip(42), device="cpu", processed_dir=GOLD,
                                 prefetch=3)
    assert_batches_equal("flip_e0", expected, list(loader))
    assert_batches_equal("flip_e1", expected, list(loader))


def test_generic_transform_callable_matches_random_flip_path(expected):
    torch.manual_seed(7)
    loader = D.create_dataloader("train", 4, True, "future", transform=O.RandomFlip(42), device="cpu", processed_dir=GOLD)
    assert_batches_equal("flip_e0", expected, list(loader))


def test_getitem_and_metadata_from_idx(expected):
    ds = D.FuturePredictionDataset("train", processed_dir=GOLD)
    for i in range(len(ds)):
        got, ref = ds[i], O.load_sample(ds.file_list[i])
        for a, b in zip(got, ref):
            assert a.dtype == b.dtype and torch.equal(a, b)
        assert repr(ds.get_metadata_from_idx(i)) == str(expected["meta_from_idx"][i])
    assert torch.equal(ds[-1][0], ds[5][0])
    with pytest.raises(IndexError):
        ds[6]
    # torch's own DataLoader over the dataset + our collate_fn is the reference arrangement verbatim
    from torch.utils.data import DataLoader
    batches = list(DataLoader(ds, batch_size=4, shuffle=False, collate_fn=lambda b: D.collate_fn(b, device="cpu")))
    assert_batches_equal("seq", expected, batches)


def test_collate_fn_empty_and_none_samples():
    out = D.collate_fn([], device="cpu")
    assert len(out) == 7 and all(t.numel() == 0 for t in out)
    ds = D.FuturePredictionDataset("train", processed_dir=GOLD)
    s = ds[0]
    got = D.collate_fn([(None,) * 6, s], device="cpu")
    ref = O.collate([(None,) * 6, s])
    assert all(torch.equal(a, b) for a, b in zip(got, ref))


# ---- fresh archives: oracle restatement vs native reader ----------------------------------------------------
@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("tiles"))
    O.write_synthetic_split(root, "train", 11, 20, 28, seed=5, t_range=(30, 47))
    O.write_synthetic_split(root, "val", 3, 20, 28, seed=6, t_range=(40, 40), compressed=False)
    return root


def test_oracle_restatement_equals_reference_golden(expected):
    files = O.list_split(GOLD, "train")
    batches = [O.collate([O.load_sample(f) for f in files[i:i + 4]]) for i in (0, 4)]
    assert_batches_equal("seq", expected, batches)


@pytest.mark.parametrize("split,batch", [("train", 4), ("train", 11), ("train", 1), ("val", 2)])
def test_native_reader_equals_oracle_on_fresh_archives(synth, split, batch):
    files = O.list_split(synth, split)
    loader = D.create_dataloader(split, batch, False, "future", device="cpu", processed_dir=synth, num_workers=3)
    got = list(loader)
    assert len(got) == -(-len(files) // batch)
    for b, g in enumerate(got):
        ref = O.collate([O.load_sample(f) for f in files[b * batch:(b + 1) * batch]])
        for a, r in zip(g, ref):
            assert a.dtype == r.dtype and torch.equal(a, r)
    st = loader.dataset.stats()
    assert st["samples"] == len(files) and st["payload_bytes"] > 0


@pytest.mark.parametrize("flip", [False, True])
def test_stage_bf16_yields_the_engines_staged_layout_bit_exactly(synth, flip):
    """create_dataloader(..., stage_bf16=True): `inputs` arrive as bf16 NHWC [B, H, W, 24] (converted inside the decode
    pass, RandomFlip included) and must equal engine.stage_maps of the fp32 batches; everything else is unchanged.  An
    arbitrary transform callable is refused (it would only see the fp32 copy)."""
    from mau_b200 import engine

    def loaders(**kw):
        torch.manual_seed(7)
        tf = D.RandomFlip(3) if flip else None
        return D.create_dataloader("train", 4, True, "future", transform=tf, device="cpu", processed_dir=synth, num_workers=3, **kw)
    plain, staged = list(loaders()), list(loaders(stage_bf16=True))
    assert len(plain) == len(staged) > 0
    for a, b in zip(plain, staged):
        want = engine.stage_maps(a[0])
        assert b[0].dtype == torch.bfloat16 and b[0].shape == want.shape == (a[0].shape[0], a[0].shape[2], a[0].shape[3], 24)
        assert torch.equal(b[0].view(torch.int16), want.view(torch.int16))
        for k in range(1, 7):
            assert torch.equal(a[k], b[k]), k
    with pytest.raises(ValueError, match="stage_bf16"):
        list(D.create_dataloader("train", 4, False, "future", transform=lambda x, y: (x, y), device="cpu", processed_dir=synth,
                                 stage_bf16=True))


def test_drop_last_and_len(synth):
    loader = D.create_dataloader("train", 4, False, "future", device="cpu", processed_dir=synth, drop_last=True)
    assert len(loader) == 2 and sum(b[0].shape[0] for b in loader) == 8


def test_rank_slices_partition_the_global_batch(synth):
    files = O.list_split(synth, "train")
    per_rank = []
    for r in range(2):
        torch.manual_seed(99)
        tf = D.RandomFlip(3)
        loader = D.create_dataloader("train", 3, True, "future", transform=tf, device="cpu", processed_dir=synth, rank=r, world_size=2)
        per_rank.append(list(loader))
    torch.manual_seed(99)
    whole = list(D.create_dataloader("train", 6, True, "future", transform=D.RandomFlip(3), device="cpu", processed_dir=synth))
    assert len(per_rank[0]) == len(per_rank[1]) == len(whole) == 2
    for b in range(2):
        for k in range(7):      # incl. the series: every rank pads to the longest series of the global batch
            assert torch.equal(torch.cat([per_rank[0][b][k], per_rank[1][b][k]]), whole[b][k]), (b, k)
    assert len(files) == 11


def test_other_dtypes_are_converted_like_tensor_float(tmp_path):
    d = tmp_path / "train"
    d.mkdir()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 4, 5))
    np.savez_compressed(d / "A_1_0.5_0.25_2019_7_to_2023_7.npz", input=x, target=(x[:2] * 1000).astype(np.int64),
                        metadata=np.arange(4, dtype=np.uint8), temperature_serie=rng.standard_normal(7).astype(np.float16))
    ds = D.FuturePredictionDataset("train", processed_dir=str(tmp_path))
    got, ref = ds[0], O.load_sample(ds.file_list[0])
    for a, b in zip(got, ref):
        assert a.dtype == torch.float32 and torch.equal(a, b)


def test_series_longer_than_the_staging_capacity_grows_the_buffer(tmp_path):
    d = tmp_path / "train"
    d.mkdir()
    rng = np.random.default_rng(2)
    for i, n in enumerate((1500, 3)):
        np.savez_compressed(d / f"B_{i}_0.5_0.25_2019_7_to_2023_7.npz", input=np.zeros((2, 2, 2), np.float32), target=np.zeros((1, 2, 2), np.float32),
                            metadata=np.zeros(4, np.float32), temperature_serie=rng.standard_normal(n).astype(np.float32))
    (batch,) = list(D.create_dataloader("train", 2, False, "future", device="cpu", processed_dir=str(tmp_path)))
    ref = O.collate([O.load_sample(f) for f in O.list_split(str(tmp_path), "train")])
    assert batch[2].shape == (2, 1500) and torch.equal(batch[2], ref[2]) and torch.equal(batch[3], ref[3])


# ---- error behaviour ---------------------------------------------------------------------------------------
def _one(tmp_path, name="C_1_0.5_0.25_2019_7_to_2023_7.npz", **arrays):
    d = tmp_path / "train"
    d.mkdir(exist_ok=True)
    base = dict(input=np.ones((2, 3, 4), np.float32), target=np.ones((1, 3, 4), np.float32), metadata=np.ones(4, np.float32),
                temperature_serie=np.ones(5, np.float32))
    base.update(arrays)
    base = {k: v for k, v in base.items() if v is not None}
    np.savez_compressed(d / name, **base)
    return str(d / name)


def test_missing_split_directory_raises_file_not_found(tmp_path):
    with pytest.raises(FileNotFoundError):
        D.FuturePredictionDataset("test", processed_dir=str(tmp_path))


def test_missing_member_raises_key_error(tmp_path):
    _one(tmp_path, temperature_serie=None)
    ds = D.FuturePredictionDataset("train", processed_dir=str(tmp_path))
    with pytest.raises(KeyError, match="temperature_serie"):
        ds[0]


def test_corrupt_archives_raise_instead_of_returning_garbage(tmp_path):
    p = _one(tmp_path, input=np.random.default_rng(0).standard_normal((2, 30, 40)).astype(np.float32))
    raw = bytearray(open(p, "rb").read())
    with zipfile.ZipFile(p) as z:
        info = z.getinfo("input.npy")
    start = info.header_offset + 30 + len(info.filename) + len(info.extra or b"") + 20        # local extra may differ; stay inside the payload
    raw[start + 400] ^= 0x55
    open(p, "wb").write(bytes(raw))
    ds = D.FuturePredictionDataset("train", processed_dir=str(tmp_path))
    with pytest.raises(ValueError):
        ds[0]
    open(p, "wb").write(bytes(raw[:len(raw) // 2]))          # truncated: no end-of-central-directory record
    with pytest.raises(ValueError):
        ds[0]
    open(p, "wb").write(b"")
    with pytest.raises(ValueError):
        ds[0]
    os.remove(p)
    with pytest.raises(FileNotFoundError):
        ds[0]


def test_crc_mismatch_is_detected_unless_disabled(tmp_path):
    p = _one(tmp_path)
    with zipfile.ZipFile(p) as z:
        info = z.getinfo("metadata.npy")
    raw = bytearray(open(p, "rb").read())
    pos = 0                 # patch the CRC field of metadata.npy in the central directory
    while True:
        pos = raw.find(b"PK\x01\x02", pos)
        assert pos >= 0
        nlen = struct.unpack_from("<H", raw, pos + 28)[0]
        if raw[pos + 46:pos + 46 + nlen] == b"metadata.npy":
            struct.pack_into("<I", raw, pos + 16, (info.CRC + 1) & 0xFFFFFFFF)
            break
        pos += 4
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="CRC"):
        D.FuturePredictionDataset("train", processed_dir=str(tmp_path))[0]
    ok = D.FuturePredictionDataset("train", processed_dir=str(tmp_path), verify_crc=False)[0]
    assert torch.equal(ok[1], torch.ones(4))


def test_shape_mismatch_inside_a_batch_raises_runtime_error(tmp_path):
    _one(tmp_path, name="C_1_0.5_0.25_2019_7_to_2023_7.npz")
    _one(tmp_path, name="C_2_0.5_0.25_2019_7_to_2023_7.npz", input=np.ones((2, 3, 5), np.float32))
    loader = D.create_dataloader("train", 2, False, "future", device="cpu", processed_dir=str(tmp_path))
    with pytest.raises(RuntimeError, match="shape"):
        list(loader)


def test_malformed_file_name_raises_like_int(tmp_path):
    _one(tmp_path, name="C_1_0.5_0.25_2019_7_to_2023_x.npz")
    with pytest.raises(ValueError):
        D.FuturePredictionDataset("train", processed_dir=str(tmp_path))[0]
    with pytest.raises(ValueError):
        O.load_sample(O.list_split(str(tmp_path), "train")[0])


def test_fortran_ordered_member_is_refused(tmp_path):
    _one(tmp_path, input=np.asfortranarray(np.arange(24, dtype=np.float32).reshape(2, 3, 4)))
    with pytest.raises(ValueError, match="Fortran"):
        D.FuturePredictionDataset("train", processed_dir=str(tmp_path))[0]


def test_async_tickets_out_of_order_and_abandoned_iteration(synth):
    ds = D.FuturePredictionDataset("train", processed_dir=synth, threads=4)
    sets = [ds.alloc_staging(3) for _ in range(3)]
    tickets = [ds.submit([3 * k, 3 * k + 1, 3 * k + 2], None, sets[k]) for k in range(3)]
    files = ds.file_list
    for k in (2, 0, 1):
        st = tickets[k].wait()
        assert tickets[k].n == 3
        ref = O.collate([O.load_sample(f) for f in files[3 * k:3 * k + 3]])
        assert torch.equal(st["input"][:3], ref[0]) and torch.equal(st["target"][:3], ref[6])
    L = D.lib()
    assert L.mau_tiles_wait(ds._handle, 12345) == D.E_ARG and b"unknown ticket" in L.mau_tiles_last_error()
    it = iter(D.TileLoader(ds, 2, False, device="cpu", prefetch=3))
    next(it)
    it.close()          # in-flight decodes are drained before their buffers are dropped
    ds.close()


# ---- the reader's own DEFLATE decoder and CRC-32 against zlib ------------------------------------------------
def _deflate(data, level=6, strategy=0, flush_every=None, memlevel=8):
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
    if not flush_every:
        return c.compress(data) + c.flush()
    out = b""
    for k, i in enumerate(range(0, len(data), flush_every)):
        out += c.compress(data[i:i + flush_every]) + c.flush(zlib.Z_FULL_FLUSH if k % 2 else zlib.Z_SYNC_FLUSH)
    return out + c.flush()


def _inflate(raw, n, split):
    import ctypes as C
    src = (C.c_uint8 * (len(raw) + 16)).from_buffer_copy(raw + b"\xAA" * 16)      # 16 readable bytes of slack
    dst = (C.c_uint8 * max(n, 1))()
    rc = D.lib().mau_tiles_inflate(src, len(raw), dst, n, split)
    return rc, bytes(dst[:n])


def _payloads():
    rng = np.random.default_rng(0)
    cls = rng.integers(0, 9, (40, 40))
    yield b""
    yield b"a"
    yield bytes(70000)                                                              # distance-1 runs, length-258 matches
    yield rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()                     # incompressible: literals / stored blocks
    yield (cls[None] == np.arange(9)[:, None, None]).astype(np.float32).tobytes()   # one-hot planes: distance-4 periods
    yield rng.standard_normal(20000).astype(np.float32).tobytes()
    yield b"0123456" * 9000 + b"abc" * 5000 + b"xy" * 7000 + b"12345" * 3000 + b"abcdef" * 3000   # periods 7, 3, 2, 5, 6
    yield open(__file__, "rb").read() * 2


def test_inflate_is_bit_identical_to_zlib_on_every_block_type():
    import zlib
    n_checked = 0
    for data in _payloads():
        for level in (0, 1, 6, 9):                    # 0 = stored blocks
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                for flush_every, memlevel in ((None, 8), (7001, 8), (None, 1)):
                    raw = _deflate(data, level, strategy, flush_every, memlevel)
                    assert zlib.decompress(raw, -15) == data
                    for split in {0, len(data), len(data) // 2, min(len(data), 65536), min(len(data), 305), max(len(data) - 1, 0)}:
                        rc, out = _inflate(raw, len(data), split)
                        assert rc == 0 and out == data, (len(data), level, strategy, flush_every, memlevel, split)
                        n_checked += 1
    assert n_checked > 1500


def test_inflate_resumes_at_every_output_offset():
    data = open(__file__, "rb").read()[:2500] + bytes(400) + b"ab" * 300 + b"xyz" * 100
    raw = _deflate(data)
    for split in range(len(data) + 1):
        rc, out = _inflate(raw, len(data), split)
        assert rc == 0 and out == data, split


def test_inflate_rejects_bad_streams_without_crashing():
    rng = random_bytes = np.random.default_rng(3)
    data = open(__file__, "rb").read()[:6000]
    raw = _deflate(data)
    assert _inflate(raw, len(data) - 1, 0)[0] == D.E_FORMAT          # more data than the directory entry says
    assert _inflate(raw, len(data) + 1, 0)[0] == D.E_FORMAT          # stream ends early
    assert _inflate(raw[:len(raw) // 2], len(data), 0)[0] == D.E_FORMAT
    assert _inflate(b"\x07", 10, 0)[0] == D.E_FORMAT                # block type 3
    wrong = 0
    for _ in range(1500):                                            # bit flips: an error or different bytes, never a crash
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = _inflate(bytes(b), len(data), int(rng.integers(0, len(data))))
        wrong += rc != 0 or out != data
    assert wrong > 1000
    for _ in range(300):
        _inflate(random_bytes.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8).tobytes(), int(rng.integers(0, 4000)), 0)


def test_crc32_equals_zlib_for_all_lengths_and_continuations():
    import zlib
    buf = np.random.default_rng(1).integers(0, 256, 1 << 18, dtype=np.uint8).tobytes()
    crc = D.lib().mau_tiles_crc32
    for n in list(range(0, 200)) + [1000, 4095, 4096, 65537, (1 << 18) - 3]:
        for off in (0, 1, 7):
            d = buf[off:off + n]
            assert crc(0, d, len(d)) == zlib.crc32(d), (n, off)
            k = len(d) // 3
            assert crc(crc(0, d[:k], k), d[k:], len(d) - k) == zlib.crc32(d)


def test_zlib_switch_decodes_the_same_batches(synth):
    a = list(D.TileLoader(D.FuturePredictionDataset("train", processed_dir=synth, use_zlib=True), 4, False, device="cpu"))
    b = list(D.TileLoader(D.FuturePredictionDataset("train", processed_dir=synth), 4, False, device="cpu"))
    assert len(a) == len(b) == 3
    assert all(torch.equal(x, y) for ba, bb in zip(a, b) for x, y in zip(ba, bb))


def test_members_larger_than_the_decoder_head_buffer(tmp_path):
    # the first 64 KiB of a member are decoded into a scratch buffer (NPY header + start of the payload), the rest
    # straight into the batch slot with the copied part as match history: 23 x 64 x 72 fp32 = 424 KB crosses that seam
    root = str(tmp_path)
    files = O.write_synthetic_split(root, "train", 5, 64, 72, seed=8, t_range=(100, 120))
    for flip in (None, D.RandomFlip(1)):
        got = list(D.create_dataloader("train", 3, False, "future", transform=flip, device="cpu", processed_dir=root, num_workers=2))
        tf = O.RandomFlip(1) if flip else None
        for b, g in enumerate(got):
            ref = O.collate([O.load_sample(f, tf) for f in files[3 * b:3 * b + 3]])
            assert all(torch.equal(a, r) for a, r in zip(g, ref))


@pytest.mark.parametrize("compressed,use_zlib,dtype", [(False, False, np.float32), (True, True, np.float32), (True, False, np.float64),
                                                       (False, False, np.float64), (True, False, np.float32)])
def test_flip_on_every_decode_path(tmp_path, compressed, use_zlib, dtype):
    # stored members (row-wise reversed copy), zlib streaming, converted dtypes and the chunked fast path all flip alike
    d = tmp_path / "train"
    d.mkdir()
    rng = np.random.default_rng(4)
    save = np.savez_compressed if compressed else np.savez
    for i in range(3):
        save(d / f"F_{i}_0.5_0.25_2019_7_to_2023_7.npz", input=rng.standard_normal((5, 130, 141)).astype(dtype),
             target=rng.standard_normal((2, 130, 141)).astype(dtype), metadata=rng.standard_normal(4).astype(np.float32),
             temperature_serie=rng.standard_normal(9).astype(np.float32))
    ds = D.FuturePredictionDataset("train", processed_dir=str(tmp_path), use_zlib=use_zlib)
    st = ds.read_batch([0, 1, 2], [True, False, True])
    files = O.list_split(str(tmp_path), "train")
    for k, flip in enumerate([True, False, True]):
        x, md, ts, t1, t2, y = O.load_sample(files[k])
        assert torch.equal(st["input"][k], x.flip(2) if flip else x)
        assert torch.equal(st["target"][k], y.flip(2) if flip else y)
        assert torch.equal(st["metadata"][k], md)


def test_repack_to_stored_members_is_lossless_and_numpy_readable(synth, tmp_path):
    import zipfile
    src, dst = os.path.join(synth, "train"), str(tmp_path / "stored" / "train")
    extra = os.path.join(src, "Extra Members_1_0.5_0.25_2019_7_to_2023_7.npz")
    np.savez_compressed(extra, input=np.ones((23, 20, 28), np.float32), target=np.zeros((2, 20, 28), np.float32), metadata=np.zeros(4, np.float32),
                        temperature_serie=np.ones(33, np.float32), note=np.arange(7), empty=np.zeros((0, 3)))
    try:
        assert D.repack_split(src, dst, workers=3) == 12
        for name in sorted(os.listdir(src)):
            with np.load(os.path.join(src, name)) as a, np.load(os.path.join(dst, name)) as b:
                assert a.files == b.files
                for k in a.files:
                    assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k])
            with zipfile.ZipFile(os.path.join(dst, name)) as z:
                assert z.testzip() is None and all(i.compress_type == zipfile.ZIP_STORED for i in z.infolist())
        a = list(D.create_dataloader("train", 5, False, "future", transform=D.RandomFlip(2), device="cpu", processed_dir=synth))
        b = list(D.create_dataloader("train", 5, False, "future", transform=D.RandomFlip(2), device="cpu", processed_dir=str(tmp_path / "stored")))
        assert len(a) == len(b) == 3 and all(torch.equal(x, y) for ba, bb in zip(a, b) for x, y in zip(ba, bb))
        assert not [f for f in os.listdir(dst) if ".tmp" in f]
    finally:
        os.remove(extra)
    with pytest.raises(FileNotFoundError):
        D.lib().mau_tiles_repack(b"/nonexistent/a.npz", os.fsencode(str(tmp_path / "x.npz"))) and D._raise(D.E_IO)


def test_dataset_survives_fork_and_pickle_in_torch_dataloader_workers(expected):
    # threads do not survive fork(): every DataLoader worker process must open its own reader
    import pickle
    from torch.utils.data import DataLoader
    ds = D.FuturePredictionDataset("train", processed_dir=GOLD, threads=2)
    ds[0]                                             # the parent's pool exists before the workers fork
    collate = lambda b: D.collate_fn(b, device="cpu")  # noqa: E731
    for ctx in ("fork", "spawn"):
        kw = dict(collate_fn=collate) if ctx == "fork" else dict(collate_fn=_collate_cpu)
        batches = list(DataLoader(ds, batch_size=4, shuffle=False, num_workers=2, multiprocessing_context=ctx, **kw))
        assert_batches_equal("seq", expected, batches)
    clone = pickle.loads(pickle.dumps(ds))
    assert torch.equal(clone[3][0], ds[3][0])


def _collate_cpu(batch):
    return D.collate_fn(batch, device="cpu")


def test_two_concurrent_iterators_over_one_loader_do_not_share_staging(synth):
    loader = D.create_dataloader("train", 3, False, "future", device="cpu", processed_dir=synth)
    ref = list(loader)
    a, b = iter(loader), iter(loader)
    got_a, got_b = [], []
    for _ in range(len(ref)):
        got_a.append(next(a))
        got_b.append(next(b))
    for g in (got_a, got_b):
        assert all(torch.equal(x, y) for bg, br in zip(g, ref) for x, y in zip(bg, br))


def test_reader_survives_corrupted_archives(tmp_path):
    """Bytes of a good archive flipped anywhere (local headers, deflate streams, NPY headers, central directory, end
    record): the reader must raise a Python exception or return data, never fault."""
    d = tmp_path / "train"
    d.mkdir()
    rng = np.random.default_rng(12)
    name = "Z_1_0.5_0.25_2019_7_to_2023_7.npz"
    good = {}
    for kind, save in (("deflate", np.savez_compressed), ("stored", np.savez)):
        save(d / name, input=rng.standard_normal((3, 40, 50)).astype(np.float32), target=rng.standard_normal((2, 40, 50)).astype(np.float32),
             metadata=rng.standard_normal(4).astype(np.float32), temperature_serie=rng.standard_normal(30).astype(np.float32))
        good[kind] = open(d / name, "rb").read()
    ds = D.FuturePredictionDataset("train", processed_dir=str(tmp_path), threads=2)
    raised = 0
    for kind, blob in good.items():
        n = len(blob)
        hot = list(range(0, 200)) + list(range(n - 600, n))            # headers and the directories at both ends
        for trial in range(400):
            b = bytearray(blob)
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.choice(hot)) if trial % 2 else int(rng.integers(0, n))
                b[pos] = int(rng.integers(0, 256))
            if trial % 50 == 0:
                b = b[:int(rng.integers(1, n))]                         # truncation
            open(d / name, "wb").write(bytes(b))
            try:
                ds.probe(0)
                ds[0]
            except (ValueError, KeyError, RuntimeError, OSError, IndexError):
                raised += 1
    assert raised > 100


def test_series_lengths_lookup_is_cached_and_reports_missing_members(synth, tmp_path):
    ds = D.FuturePredictionDataset("train", processed_dir=synth, threads=3)
    want = [int(np.load(f)["temperature_serie"].shape[0]) for f in ds.file_list]
    assert ds.series_lengths(range(len(ds))) == want
    assert ds.series_lengths([3, 3, 0]) == [want[3], want[3], want[0]]
    assert set(ds._series_len) == set(range(len(ds)))
    with pytest.raises(IndexError):
        ds.series_lengths([len(ds)])
    _one(tmp_path, temperature_serie=None)
    with pytest.raises(KeyError, match="temperature_serie"):
        D.FuturePredictionDataset("train", processed_dir=str(tmp_path)).series_lengths([0])


def test_getitem_applies_a_transform_exactly_once(expected):
    # torch DataLoader over the dataset (the reference arrangement) with the reference-style callable and with ours
    from torch.utils.data import DataLoader
    for make in (lambda: O.RandomFlip(42), lambda: D.RandomFlip(42)):
        torch.manual_seed(7)
        ds = D.FuturePredictionDataset("train", transform=make(), processed_dir=GOLD)
        batches = list(DataLoader(ds, batch_size=4, shuffle=True, collate_fn=_collate_cpu))
        assert_batches_equal("flip_e0", expected, batches)


def test_empty_split_yields_nothing(tmp_path):
    (tmp_path / "val").mkdir()
    loader = D.create_dataloader("val", 4, False, "future", device="cpu", processed_dir=str(tmp_path))
    assert len(loader) == 0 and list(loader) == []
