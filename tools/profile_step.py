"""torch.profiler breakdown of one LMD16 training step (host-side PyTorch ops vs our kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyphemus_b200 as pb
from polyphemus_b200.train import TrainStep, device_batch, synthetic_host_batch
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda", 0)
pb.set_precision(prec)
torch.manual_seed(0)
model = pb.VAE(**bench.MODEL_CFG, device=dev).to(dev).train()
step = TrainStep(model, autocast_bf16=prec == "bf16", **bench.ADAM)
host = synthetic_host_batch(256, 16, 0.25, seed=0)
for _ in range(3):
    step(device_batch(host, dev))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step(device_batch(host, dev))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print(torch.cuda.max_memory_allocated() / 2**30, "GiB peak")
