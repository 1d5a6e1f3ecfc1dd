"""Pure host cost of one training step: the same step on a tiny batch (device work negligible), wall clock per step."""
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
for m in model.modules():
    if isinstance(m, pb.GCL): m.dropout = 0.1
step = TrainStep(model, autocast_bf16=True, **bench.ADAM)
host = synthetic_host_batch(2, 16, 0.25, seed=0)
for i in range(5): step(device_batch(host, dev))
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20): step(device_batch(host, dev))
torch.cuda.synchronize()
print(f"tiny batch (2 sequences): {1e3*(time.perf_counter()-t0)/20:.2f} ms/step wall = host floor of the step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(10): step(device_batch(host, dev))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
