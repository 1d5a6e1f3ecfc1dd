"""Host (enqueue) time vs device time of the training step: is the step launch-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import polyphemus_b200 as pb
from polyphemus_b200.train import TrainStep, device_batch, synthetic_host_batch
dev = torch.device("cuda", 0)
pb.set_precision("bf16")
torch.manual_seed(0)
model = pb.VAE(**bench.MODEL_CFG, device=dev).to(dev).train()
step = TrainStep(model, autocast_bf16=True, **bench.ADAM)
host = synthetic_host_batch(256, 16, 0.25, seed=0)
graphs = [device_batch(host, dev) for _ in range(12)]
for g in graphs: g.structured  # plans prebuilt
for i in range(3): step(graphs[i])
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3, 11): step(graphs[i])
t1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"8 steps, graphs prebuilt: host enqueue {1e3*(t1-t0)/8:.2f} ms/step, device {e0.elapsed_time(e1)/8:.2f} ms/step, wall {1e3*(t2-t0)/8:.2f}")
# with graph build inline
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(8): step(device_batch(host, dev))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"8 steps, inline graph build: host enqueue {1e3*(t1-t0)/8:.2f} ms/step, wall {1e3*(t2-t0)/8:.2f}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(4): step(graphs[i])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
print("---- sync debug: every host sync inside one step")
torch.cuda.set_sync_debug_mode("warn")
import warnings
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    step(graphs[0])
    g2 = device_batch(host, dev)
torch.cuda.set_sync_debug_mode("default")
for x in w:
    print("SYNC:", str(x.message)[:100], "@", x.filename.split("/")[-1], x.lineno)
print("---- host time per ABI entry point (perf_counter around the ctypes call), 4 steps")
import collections
from polyphemus_b200 import _ffi as F
acc = collections.defaultdict(lambda: [0, 0.0])
orig = F.call
def timed_call(name, *a, tag=None):
    t = time.perf_counter()
    orig(name, *a, tag=tag)
    e = acc[name]; e[0] += 1; e[1] += time.perf_counter() - t
F.call = timed_call
torch.cuda.synchronize()
for i in range(4):
    step(device_batch(host, dev)); torch.cuda.synchronize()   # synchronised: no launch-queue back-pressure in the numbers
F.call = orig
tot = sum(v[1] for v in acc.values())
print(f"total {1e3*tot/4:.2f} ms/step in {sum(v[0] for v in acc.values())//4} calls/step")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:40s} {v[0]//4:4d} calls/step  {1e6*v[1]/v[0]:7.1f} us/call  {1e3*v[1]/4:6.2f} ms/step")
