"""GPU: the tile loader feeding the hot path -- pinned staging ring + side-stream H2D copies deliver exactly the
batches of the CPU oracle (oracle/dataset_oracle.py), also when the consumer synchronises every step."""
import numpy as np
import pytest
import torch

import mau_b200
from mau_b200 import data as D
from oracle import dataset_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("tiles_gpu"))
    O.write_synthetic_split(root, "train", 13, 48, 48, seed=21, t_range=(30, 40))
    return root


def two_epochs(root, device, prefetch, sync_each_step=False):
    torch.manual_seed(5)
    loader = D.create_dataloader("train", 4, True, "future", transform=D.RandomFlip(9), device=device, processed_dir=root,
                                 prefetch=prefetch)
    got = []
    for _ in range(2):
        for batch in loader:
            if sync_each_step:
                torch.cuda.synchronize()
            got.append(batch)
    return got


@pytest.mark.parametrize("prefetch,sync_each_step", [(1, True), (2, False), (4, True)])
def test_device_batches_equal_the_cpu_loader_and_oracle(synth, prefetch, sync_each_step):
    # the CPU loader is pinned bit-exact to the reference loader's golden batches in tests/test_tiles_cpu.py
    ref = two_epochs(synth, "cpu", 2)
    got = two_epochs(synth, "cuda:0", prefetch, sync_each_step)
    assert len(got) == len(ref) == 8
    for g, r in zip(got, ref):
        assert all(t.is_cuda for k, t in enumerate(g) if k != 3) and not g[3].is_cuda
        for a, b in zip(g, r):
            assert a.dtype == b.dtype and torch.equal(a.cpu(), b)
    # sequential batches against the NumPy oracle directly
    files = O.list_split(synth, "train")
    for b, g in enumerate(D.create_dataloader("train", 5, False, "future", device="cuda:0", processed_dir=synth)):
        want = O.collate([O.load_sample(f) for f in files[5 * b:5 * b + 5]])
        assert all(torch.equal(a.cpu(), w) for a, w in zip(g, want))


def test_training_steps_fed_from_archives(synth):
    torch.manual_seed(0)
    kw = dict(temporal_embeddings=True, metadata_embeddings=True)
    model = mau_b200.UrbanPredictor("unet", 23, 828, 16, 8, 16, 32, 2, base_filters=16, **kw).to("cuda:0").train()
    opt = mau_b200.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-3)
    loader = D.create_dataloader("train", 4, False, "future", device="cuda:0", processed_dir=synth, drop_last=True)
    losses = []
    for inputs, metadatas, series, lengths, t1, t2, targets in loader:
        md = torch.cat([metadatas, t1, t2], dim=1)                      # src/train.py:244
        out = model(inputs, series, md)
        loss = (out - targets).abs().mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(loss.item())                                       # src/train.py:258: one sync per step
    assert len(losses) == 3 and all(np.isfinite(losses))


def test_staged_bf16_batches_drive_the_model_to_identical_outputs(synth):
    """create_dataloader(..., stage_bf16=True) ships the tiles as bf16 NHWC (half the PCIe bytes); the model must return
    exactly what it returns for the fp32 NCHW batches of the plain loader (eval forward, and the loss of a training step)."""
    torch.manual_seed(0)
    kw = dict(temporal_embeddings=True, metadata_embeddings=True)
    model = mau_b200.UrbanPredictor("unet", 23, 828, 16, 8, 16, 32, 2, base_filters=16, **kw).to("cuda:0")

    def batches(**opt):
        torch.manual_seed(5)
        return list(D.create_dataloader("train", 4, True, "future", transform=D.RandomFlip(9), device="cuda:0", processed_dir=synth,
                                        drop_last=True, **opt))
    plain, staged = batches(), batches(stage_bf16=True)
    assert len(plain) == len(staged) == 3
    for p, s in zip(plain, staged):
        assert s[0].dtype == torch.bfloat16 and s[0].shape == (4, 48, 48, 24) and s[0].is_cuda
        md = torch.cat([p[1], p[4], p[5]], dim=1)
        model.eval()
        with torch.no_grad():
            assert torch.equal(model(p[0], p[2], md), model(s[0], s[2], md))
        model.train()
        sd0 = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "tracked" in k}
        la = float((model(p[0], p[2], md) - p[6]).abs().mean())
        for k, v in sd0.items():
            model.state_dict()[k].copy_(v)
        lb = float((model(s[0], s[2], md) - s[6]).abs().mean())
        assert la == lb
