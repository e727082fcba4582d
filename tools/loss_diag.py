"""Why is a pipelined training loop slower with a torch-op loss than with the fused loss kernel? (debug helper)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import mau_b200
from mau_b200 import engine
from oracle import unet_oracle as O
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = mau_b200.UrbanPredictor("unet", 23, 828, 64, 8, 64, 96, 2, temporal_embeddings=False, metadata_embeddings=True).to(dev).train()
xd, td, mdd, tg = [t.to(dev) for t in O.synthetic_batch(16, 250, 250, seed=1)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(name, lossfn, n=10):
    for _ in range(2):
        lossfn(m(xd, td, mdd)).backward(); m.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        lossfn(m(xd, td, mdd)).backward(); m.zero_grad(set_to_none=True)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print(f"{name:34s} device {e0.elapsed_time(e1)/n:7.3f} ms/step  host issue {(t1-t0)/n*1e3:7.3f} ms/step  "
          f"cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}  frees {s1['num_device_free']-s0['num_device_free']}  "
          f"reserved {s1['reserved_bytes.all.current']/1e9:.2f} GB", flush=True)
run("engine L1 kernel", lambda o: engine.compute_loss_l1_grad(o, tg, 0.0)["total"])
run("torch out.mean()", lambda o: o.mean())
run("torch (out-tg).abs().mean()", lambda o: (o - tg).abs().mean())
run("torch F.l1_loss", lambda o: F.l1_loss(o, tg))
run("torch F.mse_loss", lambda o: F.mse_loss(o, tg))
run("engine L1 kernel (again)", lambda o: engine.compute_loss_l1_grad(o, tg, 0.0)["total"])
