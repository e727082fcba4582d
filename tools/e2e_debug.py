"""Where does the time of a training e2e step go? (debug helper)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mau_b200
from mau_b200 import engine
from oracle import unet_oracle as O
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = mau_b200.UrbanPredictor("unet", 23, 828, 64, 8, 64, 96, 2, temporal_embeddings=False, metadata_embeddings=True).to(dev).train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
B = 16
x, ts, md, tgt = [t.pin_memory() for t in O.synthetic_batch(B, 250, 250, seed=1)]
xd, td, mdd, tg = [t.to(dev) for t in (x, ts, md, tgt)]
def sync(): torch.cuda.synchronize()
def timeit(name, fn, n=5):
    fn(); sync()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    sync()
    print(f"{name:40s} {(time.perf_counter()-t0)/n*1e3:8.2f} ms", flush=True)
def step_dev():
    out = m(xd, td, mdd); loss = engine.compute_loss_l1_grad(out, tg, 0.0)["total"]; loss.backward(); opt.zero_grad(set_to_none=True)
def step_h2d():
    a, b, c, d = [t.to(dev, non_blocking=True) for t in (x, ts, md, tgt)]
    out = m(a, b, c); loss = engine.compute_loss_l1_grad(out, d, 0.0)["total"]; loss.backward(); opt.zero_grad(set_to_none=True)
    return loss
def h2d_only():
    return [t.to(dev, non_blocking=True) for t in (x, ts, md, tgt)]
copy_s = torch.cuda.Stream()
def step_prefetch():
    with torch.cuda.stream(copy_s):
        a, b, c, d = [t.to(dev, non_blocking=True) for t in (x, ts, md, tgt)]
        ev = torch.cuda.Event(); ev.record(copy_s)
    torch.cuda.current_stream().wait_event(ev)
    for t in (a, b, c, d): t.record_stream(torch.cuda.current_stream())
    out = m(a, b, c); loss = engine.compute_loss_l1_grad(out, d, 0.0)["total"]; loss.backward(); opt.zero_grad(set_to_none=True)
    return loss.detach().cpu()
timeit("train step, device inputs", step_dev)
timeit("h2d only (100 MB)", h2d_only)
timeit("train step + h2d same stream", step_h2d)
timeit("train step + h2d side stream + .cpu()", step_prefetch)
def fwd_only():
    with torch.no_grad(): m(xd, td, mdd)
timeit("train-mode forward only", fwd_only)
m.eval()
timeit("eval forward only", fwd_only)
