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

# ---- where the GPU idles: gaps between consecutive device activities of the profiled steps
try:
    from torch.autograd import DeviceType
    evs = [e for e in prof.events() if e.device_type == DeviceType.CUDA and e.time_range.end > e.time_range.start]
    evs.sort(key=lambda e: e.time_range.start)
    gaps, busy_end = [], None
    for prev, cur in zip(evs, evs[1:]):
        busy_end = prev.time_range.end if busy_end is None else max(busy_end, prev.time_range.end)
        gap = cur.time_range.start - busy_end
        if gap > 0:
            gaps.append((gap, prev.name[:60], cur.name[:60]))
    total = sum(g for g, _, _ in gaps)
    span = evs[-1].time_range.end - evs[0].time_range.start
    print(f"device span {span / 1e3:.2f} ms for 2 steps, idle {total / 1e3:.2f} ms in {len(gaps)} gaps")
    import collections
    by_next = collections.defaultdict(lambda: [0, 0.0])
    for g, p, c in gaps:
        by_next[c][0] += 1
        by_next[c][1] += g
    for name, (n, t) in sorted(by_next.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  idle before {name:62s} x{n:4d}  {t / 1e3:7.3f} ms")
    for g, p, c in sorted(gaps, reverse=True)[:12]:
        print(f"  gap {g:8.1f} us  after {p}  before {c}")
except Exception as exc:       # profiler API drift: the table above is the primary output
    print("gap analysis unavailable:", exc)
