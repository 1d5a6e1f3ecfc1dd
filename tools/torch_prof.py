"""Attribute the non-library device time of one training step to the torch ops that launch it (torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import polyphemus_b200 as pb
from polyphemus_b200.train import TrainStep, device_batch, synthetic_host_batch
dev = torch.device("cuda", 0)
pb.set_precision("bf16")
torch.manual_seed(0)
model = pb.VAE(**bench.MODEL_CFG, device=dev).to(dev).train()
step = TrainStep(model, autocast_bf16=True, **bench.ADAM)
host = synthetic_host_batch(256, 16, 0.25, seed=0)
for i in range(3): step(device_batch(host, dev))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    for i in range(2): step(device_batch(host, dev))
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = sorted(ka, key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"total self device time {tot/2e3:.2f} ms/step")
for e in rows[:70]:
    print(f"{e.self_device_time_total/2:9.1f} us/step  n={e.count//2:4d}  {e.key[:60]:60s} {str(e.input_shapes)[:110]}")
print("---- by stack (top frames in polyphemus_b200) for aten ops")
ks = prof.key_averages(group_by_stack_n=12)
agg = {}
for e in ks:
    if e.self_device_time_total <= 0 or not e.key.startswith("aten::"): continue
    fr = [f for f in e.stack if "polyphemus_b200" in f or "bench.py" in f]
    key = (e.key, fr[0].split("polyphemus_b200/")[-1][:70] if fr else "?")
    agg[key] = agg.get(key, 0) + e.self_device_time_total
for (k, f), v in sorted(agg.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{v/2:9.1f} us/step  {k:34s} {f}")
