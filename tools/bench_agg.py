"""Micro-benchmark of the aggregation kernels on the bench graph (LMD16, batch 256, d = 512, bf16, structured layout):
CUDA-event time per launch of pb_agg_fwd, pb_agg_bwd (legacy) and pb_agg_bwd_fused, inputs larger than L2.
    python tools/bench_agg.py [--iters 20] [--d 512] [--batch 256]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyphemus_b200 as pb
from polyphemus_b200 import _ffi
from polyphemus_b200.train import synthetic_host_batch

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--d", type=int, default=512)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--which", default="fwd,bwd_legacy,bwd_fused")
args = ap.parse_args()
dev = torch.device("cuda", 0)
host = synthetic_host_batch(args.batch, 16, 0.25, seed=0, pin=False)
g = pb.graphs_from_tensor(host.s_tensor.to(dev))
stp = g.structured
plan, n, d = stp.plan, stp.n_padded, args.d
k, e = 4 * d, g.num_edges
lib = _ffi.lib()
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(n, d, device=dev).to(torch.bfloat16)
gy = torch.randn(n, d, device=dev).to(torch.bfloat16)
d_a = torch.randn(n, k, device=dev).to(torch.bfloat16)
a_op = torch.empty(n, k, device=dev, dtype=torch.bfloat16)
table = torch.randn(32, d, device=dev) * 0.5
bits = torch.empty(lib.pb_dropout_bits_bytes(e, d) // 2, dtype=torch.int16, device=dev)
_ffi.check(lib.pb_dropout_bits(e, d, 0.1, 1234, bits.data_ptr(), st))
gx = torch.empty(n, d, device=dev, dtype=torch.bfloat16)
q_buf = torch.empty(e, d, device=dev, dtype=torch.bfloat16)
parts = torch.empty(plan.n_dist_items, d, device=dev)
n_part = lib.pb_agg_bwd_num_partials(n, d, _ffi.PB_BF16)
fparts = torch.empty(n_part, 32, d, device=dev)
gw, gb = torch.empty(d, 32, device=dev), torch.empty(d, device=dev)
calls = {
    "fwd": lambda: lib.pb_agg_fwd(plan.ref(), x.data_ptr(), d, table.data_ptr(), a_op.data_ptr(), None, k, _ffi.PB_BF16, bits.data_ptr(), 0.1, _ffi.PB_BF16, st),
    "bwd_legacy": lambda: (lib.pb_agg_bwd(plan.ref(), x.data_ptr(), d, table.data_ptr(), d_a.data_ptr(), k, _ffi.PB_BF16, gy.data_ptr(), gx.data_ptr(), q_buf.data_ptr(), parts.data_ptr(), bits.data_ptr(), 0.1, _ffi.PB_BF16, st),
                           lib.pb_edge_table_bwd(parts.data_ptr(), plan.dist_item_ptr.data_ptr(), d, gw.data_ptr(), gb.data_ptr(), st))[0],
    "bwd_fused": lambda: (lib.pb_agg_bwd_fused(plan.ref(), x.data_ptr(), d, table.data_ptr(), d_a.data_ptr(), k, _ffi.PB_BF16, gy.data_ptr(), gx.data_ptr(), fparts.data_ptr(), bits.data_ptr(), 0.1, _ffi.PB_BF16, st),
                          lib.pb_edge_table_bwd_fused(fparts.data_ptr(), n_part, d, gw.data_ptr(), gb.data_ptr(), st))[0],
}
out = {"lib": os.path.basename(_ffi.LIB_PATH), "nodes": g.num_nodes, "rows": n, "edges": e, "d": d}
for name in args.which.split(","):
    fn = calls[name]
    for _ in range(3):
        _ffi.check(fn(), name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[name + "_us"] = round(e0.elapsed_time(e1) / args.iters * 1e3, 1)
print(json.dumps(out))
