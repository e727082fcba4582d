"""Host-side issue time vs device time of one training step (is the step launch-bound?)."""
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
B = int(os.environ.get("B", "16"))
xd, td, mdd, tg = [t.to(dev) for t in O.synthetic_batch(B, 250, 250, seed=1)]
def sync(): torch.cuda.synchronize()
for _ in range(3):
    out = m(xd, td, mdd); loss = engine.compute_loss_l1_grad(out, tg, 0.0)["total"]; loss.backward(); m.zero_grad(set_to_none=True)
sync()
N = 5
acc = {"fwd": 0.0, "loss": 0.0, "bwd": 0.0, "total_host": 0.0, "total_dev": 0.0}
for _ in range(N):
    sync(); t0 = time.perf_counter()
    out = m(xd, td, mdd); t1 = time.perf_counter()
    loss = engine.compute_loss_l1_grad(out, tg, 0.0)["total"]; t2 = time.perf_counter()
    loss.backward(); t3 = time.perf_counter()
    m.zero_grad(set_to_none=True); t4 = time.perf_counter()
    sync(); t5 = time.perf_counter()
    acc["fwd"] += t1 - t0; acc["loss"] += t2 - t1; acc["bwd"] += t3 - t2; acc["total_host"] += t4 - t0; acc["total_dev"] += t5 - t0
for k, v in acc.items():
    print(f"{k:12s} {v / N * 1e3:8.3f} ms", flush=True)
# pieces of the python shim
net = m.model
t0 = time.perf_counter()
for _ in range(20): st = net._state_tensors()
print(f"state_dict walk        {(time.perf_counter()-t0)/20*1e3:8.3f} ms")
plan = next(iter(net._plans.values()))
t0 = time.perf_counter()
for _ in range(20): plan._check_state(st)
print(f"_check_state           {(time.perf_counter()-t0)/20*1e3:8.3f} ms")
t0 = time.perf_counter()
for _ in range(20): plan._ptr_array(st)
print(f"_ptr_array             {(time.perf_counter()-t0)/20*1e3:8.3f} ms")
# raw C forward/backward issue time
import ctypes as C
L = engine.lib()
arr = plan._ptr_array(st)
outb = torch.empty(plan.out_shape, device=dev)
sync(); t0 = time.perf_counter()
for _ in range(5):
    L.mau_plan_forward(plan._h, arr, xd.data_ptr(), None, mdd.data_ptr(), outb.data_ptr(), None)
t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
print(f"C forward issue        {(t1-t0)/5*1e3:8.3f} ms   (device-complete {(t2-t0)/5*1e3:8.3f} ms)")
grads = [torch.empty_like(t) if (r == 0 and t.dtype == torch.float32) else None for t, r in zip(st, plan.roles)]
garr = plan._ptr_array(grads)
gout = torch.randn(plan.out_shape, device=dev)
L.mau_plan_forward(plan._h, arr, xd.data_ptr(), None, mdd.data_ptr(), outb.data_ptr(), None)
sync(); t0 = time.perf_counter()
rc = L.mau_plan_backward(plan._h, gout.data_ptr(), garr, None)
t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
print(f"C backward issue       {(t1-t0)*1e3:8.3f} ms   (device-complete {(t2-t0)*1e3:8.3f} ms) rc={rc}")
# device-side duration of each phase (CUDA events on the stream; includes idle gaps)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
tot = [0.0] * 4
for it in range(6):
    ev[0].record()
    out = m(xd, td, mdd); ev[1].record()
    loss = engine.compute_loss_l1_grad(out, tg, 0.0)["total"]; ev[2].record()
    loss.backward(); ev[3].record()
    m.zero_grad(set_to_none=True); ev[4].record()
    sync()
    if it:
        for k in range(4): tot[k] += ev[k].elapsed_time(ev[k + 1])
print("device ms  fwd %.3f  loss %.3f  bwd %.3f  zero_grad %.3f" % tuple(t / 5 for t in tot))
# no sync between steps (steady state)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    out = m(xd, td, mdd); loss = engine.compute_loss_l1_grad(out, tg, 0.0)["total"]; loss.backward(); m.zero_grad(set_to_none=True)
e1.record(); sync()
print("steady-state step %.3f ms" % (e0.elapsed_time(e1) / 10))
e0.record()
for it in range(10):
    out = m(xd, td, mdd); loss = (out - tg).abs().mean(); loss.backward(); m.zero_grad(set_to_none=True)
e1.record(); sync()
print("steady-state step, torch L1 loss %.3f ms" % (e0.elapsed_time(e1) / 10))
# torch-op loss (what the reference's own training loop does): per-phase device time
tot = [0.0] * 4
for it in range(6):
    ev[0].record()
    out = m(xd, td, mdd); ev[1].record()
    loss = (out - tg).abs().mean(); ev[2].record()
    loss.backward(); ev[3].record()
    m.zero_grad(set_to_none=True); ev[4].record()
    sync()
    if it:
        for k in range(4): tot[k] += ev[k].elapsed_time(ev[k + 1])
print("torch L1 loss: device ms  fwd %.3f  loss %.3f  bwd %.3f  zero_grad %.3f" % tuple(t / 5 for t in tot))
t0 = time.perf_counter(); out = m(xd, td, mdd); loss = (out - tg).abs().mean(); t1 = time.perf_counter(); loss.backward(); t2 = time.perf_counter(); sync(); t3 = time.perf_counter()
print("torch L1 loss: host ms fwd+loss %.3f  backward() call %.3f  sync %.3f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
