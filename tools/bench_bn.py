"""Micro-benchmark of the BatchNorm kernels on the bench shape (structured layout of LMD16 batch 256, d = 512, bf16)."""
import argparse, ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyphemus_b200 as pb
from polyphemus_b200 import _ffi
from polyphemus_b200.train import synthetic_host_batch
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--d", type=int, default=512)
args = ap.parse_args()
dev = torch.device("cuda", 0)
host = synthetic_host_batch(256, 16, 0.25, seed=0, pin=False)
g = pb.graphs_from_tensor(host.s_tensor.to(dev))
stp = g.structured
n, d = stp.n_padded, args.d
lib = _ffi.lib(); st = torch.cuda.current_stream().cuda_stream
bf = torch.bfloat16
out = torch.randn(n, d, device=dev).to(bf); x = torch.randn(n, d, device=dev).to(bf); y = torch.empty_like(x)
gy = torch.randn(n, d, device=dev).to(bf); g_hi = torch.empty_like(x)
gamma = torch.rand(d, device=dev) + 0.5; beta = torch.randn(d, device=dev)
rm_, rv_ = torch.zeros(d, device=dev), torch.ones(d, device=dev)
save = torch.empty(2, d, device=dev); coef = torch.empty(3, d, device=dev)
ws_bytes = lib.pb_bn_workspace_bytes(n, d); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
gg, gb, gbias = (torch.empty(d, device=dev) for _ in range(3))
gr = stp.groups_ref()
calls = {
  "bn_stats": lambda: lib.pb_bn_stats(out.data_ptr(), d, n, d, gr, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, rm_.data_ptr(), rv_.data_ptr(), save.data_ptr(), coef.data_ptr(), ws.data_ptr(), ws_bytes, _ffi.PB_BF16, st),
  "bn_fwd": lambda: lib.pb_bn_relu_res_fwd(out.data_ptr(), d, x.data_ptr(), coef.data_ptr(), y.data_ptr(), n, d, gr, 1, _ffi.PB_BF16, st),
  "bn_bwd": lambda: lib.pb_bn_relu_res_bwd(gy.data_ptr(), out.data_ptr(), d, gamma.data_ptr(), save.data_ptr(), coef.data_ptr(), n, d, gr, _ffi.PB_BF16, g_hi.data_ptr(), None, d, gg.data_ptr(), gb.data_ptr(), gbias.data_ptr(), ws.data_ptr(), ws_bytes, _ffi.PB_BF16, st),
}
res = {"lib": os.path.basename(_ffi.LIB_PATH), "rows": n, "d": d}
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # L2 flush between launches
for name, fn in calls.items():
    for _ in range(3): _ffi.check(fn(), name)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(args.iters):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    res[name + "_us"] = round(tot / args.iters * 1e3, 1)
print(json.dumps(res))
