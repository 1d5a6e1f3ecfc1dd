"""GPU diagnostic: per-tile error of the bf16 forward GEMM over a sweep of (m, d, k)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from polyphemus_b200 import _ffi as ffi
cuda = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream
def run(m, d, k, dtype=ffi.PB_BF16):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(m, k, generator=g).to(cuda); wt = (torch.randn(d, k, generator=g)/np.sqrt(k)).to(cuda)
    a_b, w_b = a.to(torch.bfloat16).contiguous(), wt.to(torch.bfloat16).contiguous()
    out = torch.full((m, d), float("nan"), device=cuda)
    ffi.check(ffi.lib().pb_rgcn_gemm_fwd(a_b.data_ptr(), None, k, w_b.data_ptr(), None, None, out.data_ptr(), d, m, d, k, dtype, st()), "fwd")
    torch.cuda.synchronize()
    ref = a_b.double() @ w_b.double().t()
    err = (out.double() - ref).abs()
    tiles = []
    for mt in range(0, m, 128):
        for nt in range(0, d, 256):
            tiles.append(f"{err[mt:mt+128, nt:nt+256].max().item():.1e}")
    print(f"m={m} d={d} k={k}: max err {err.max().item():.3e}  nan={int(torch.isnan(out).sum())}  tiles={tiles[:12]}")
for (m, d, k) in [(128, 512, 512), (128, 512, 1024), (128, 512, 2048), (128, 512, 3584), (256, 512, 3584), (1000, 512, 3584),
                  (1000, 256, 3584), (1000, 512, 1792), (1000, 384, 3584), (2000, 512, 3584), (20000, 512, 3584)]:
    run(m, d, k)
