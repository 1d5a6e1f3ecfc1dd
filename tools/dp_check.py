"""2-GPU check of the data-parallel step: NCCL-averaged gradients == average of the per-shard gradients computed
on one GPU (same weights, GCL dropout off so that both sides are deterministic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import polyphemus_b200 as pb
from polyphemus_b200.train import GradAllReducer, device_batch, synthetic_host_batch, vae_losses

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
cfg = dict(dropout=0, batch_norm=True, gnn_n_layers=2, d=128, n_bars=2, resolution=8)
pb.set_precision("fp32")
torch.manual_seed(0)
model = pb.VAE(**cfg, device=dev).to(dev).train()
for m in model.modules():
    if isinstance(m, pb.GCL):
        m.dropout = 0.0
red = GradAllReducer(model.parameters(), bucket_mb=0.05)
hosts = [synthetic_host_batch(6, 2, 0.25, seed=100 + r, pin=False) for r in range(world)]
noise = [torch.randn(6, cfg["d"], generator=torch.Generator().manual_seed(r)).to(dev) for r in range(world)]

def grads_for(shard):
    graph = device_batch(hosts[shard], dev)
    (s_logits, c_logits), mu, log_var = model(graph, noise=noise[shard])
    loss, _ = vae_losses(graph.s_tensor, s_logits, None, c_logits, mu, log_var, c_tokens=graph.c_tokens)
    loss.backward()

sd0 = {k: v.clone() for k, v in model.state_dict().items()}
red.zero_grad()
grads_for(rank)
red.finish()
got = red.flat.clone()
# single-GPU reference on this rank: run every shard with the same starting state, average
ref = torch.zeros_like(got)
red.sync = False
for shard in range(world):
    model.load_state_dict(sd0)
    red.zero_grad()
    grads_for(shard)
    red.finish()                       # sync off: packs the local gradients into the buffer, no exchange
    ref += red.flat
red.sync = True
ref /= world
err = (got - ref).abs().max().item()
scale = ref.abs().max().item()
print(f"rank {rank}: max |dp - mean(shards)| = {err:.3e} (grad scale {scale:.3e}), buckets={len(red.buckets)}")
assert err <= 1e-5 * max(1.0, scale) + 1e-6, err
dist.barrier()
if rank == 0:
    print("dp_check ok")
dist.destroy_process_group()
