"""Per-kernel GPU parity through the C ABI: tcgen05 GEMMs, aggregation fwd/bwd, BatchNorm fwd/bwd.

Checker = oracle.model_oracle (CPU, fp64/fp32) and exact fp64 matmuls of the operands; tolerances:
  PB_F32 mode  rtol 1e-4 / atol 1e-5  (BASELINE.json north_star)
  PB_BF16 mode compared against fp64 arithmetic on the *bf16-rounded operands* at rtol 2e-3 (accumulation
  only), and against the fp32 oracle at the bf16 budget rtol 3e-2 / atol 3e-2 of the output scale.
"""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

F32_TOL = dict(rtol=1e-4, atol=1e-5)


def _ffi():
    from polyphemus_b200 import _ffi

    return _ffi


def tf32_split(t):
    hi = (t.contiguous().view(torch.int32) & -8192).view(torch.float32)   # 0xFFFFE000
    return hi, t - hi


def operands(t, dtype):
    ffi = _ffi()
    if dtype == ffi.PB_BF16:
        return t.to(torch.bfloat16).contiguous(), None
    hi, lo = tf32_split(t)
    return hi.contiguous(), lo.contiguous()


def as_f64(hi, lo):
    return hi.double() if lo is None else hi.double() + lo.double()


def st():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def keep_bits(n_edges, d, p_drop, seed, cuda):
    """Packed dropout keep-bits (pb_dropout_bits) or None when dropout is off."""
    ffi = _ffi()
    if not p_drop:
        return None
    bits = torch.empty(ffi.lib().pb_dropout_bits_bytes(n_edges, d) // 2, dtype=torch.int16, device=cuda)
    ffi.check(ffi.lib().pb_dropout_bits(n_edges, d, p_drop, seed, ptr(bits), st()), "pb_dropout_bits")
    return bits


# ------------------------------------------------------------------------------------------------- GEMMs
@pytest.mark.parametrize("dtype_name", ["bf16", "fp32"])
@pytest.mark.parametrize("m,d", [(128, 64), (300, 128), (1000, 512), (4133, 256)])
def test_gemm_fwd(cuda, dtype_name, m, d):
    ffi = _ffi()
    dtype = ffi.PB_BF16 if dtype_name == "bf16" else ffi.PB_F32
    k = 7 * d
    g = torch.Generator(device="cpu").manual_seed(m + d)
    a = torch.randn(m, k, generator=g).to(cuda)
    wt = (torch.randn(d, k, generator=g) / np.sqrt(k)).to(cuda)
    bias = torch.randn(d, generator=g).to(cuda)
    a_hi, a_lo = operands(a, dtype)
    w_hi, w_lo = operands(wt, dtype)
    out = torch.full((m, d), float("nan"), device=cuda)
    ffi.check(ffi.lib().pb_rgcn_gemm_fwd(ptr(a_hi), ptr(a_lo), k, ptr(w_hi), ptr(w_lo), ptr(bias), ptr(out), d, m, d, k,
                                         None, dtype, ffi.PB_F32, st()), "pb_rgcn_gemm_fwd")
    torch.cuda.synchronize()
    ref = as_f64(a_hi, a_lo) @ as_f64(w_hi, w_lo).t() + bias.double()
    tol = dict(rtol=2e-3, atol=2e-3) if dtype_name == "bf16" else F32_TOL
    torch.testing.assert_close(out.double(), ref, **tol)
    # cross-check with the on-device fp32 CUDA-core contraction
    chk = torch.empty_like(out)
    a32, w32 = as_f64(a_hi, a_lo).float().contiguous(), as_f64(w_hi, w_lo).float().contiguous()   # keep alive
    ffi.check(ffi.lib().pb_gemm_f32_check(ptr(a32), k, ptr(w32), k, ptr(bias), ptr(chk), d, m, d, k, 0, 1, st()),
              "pb_gemm_f32_check")
    torch.cuda.synchronize()
    torch.testing.assert_close(out, chk, rtol=2e-3 if dtype_name == "bf16" else 1e-4, atol=2e-3 if dtype_name == "bf16" else 1e-5)


@pytest.mark.parametrize("dtype_name", ["bf16", "fp32"])
@pytest.mark.parametrize("m,d", [(128, 64), (777, 128), (2000, 512)])
def test_gemm_bwd_data(cuda, dtype_name, m, d):
    ffi = _ffi()
    dtype = ffi.PB_BF16 if dtype_name == "bf16" else ffi.PB_F32
    k = 7 * d
    gen = torch.Generator(device="cpu").manual_seed(7 * m + d)
    g = torch.randn(m, d, generator=gen).to(cuda)
    w = (torch.randn(k, d, generator=gen) / np.sqrt(d)).to(cuda)
    g_hi, g_lo = operands(g, dtype)
    w_hi, w_lo = operands(w, dtype)
    d_a = torch.empty((m, k), dtype=torch.bfloat16 if dtype == ffi.PB_BF16 else torch.float32, device=cuda)
    ffi.check(ffi.lib().pb_rgcn_gemm_bwd_data(ptr(g_hi), ptr(g_lo), d, ptr(w_hi), ptr(w_lo), ptr(d_a), k, m, d, k, None, dtype,
                                              st()), "pb_rgcn_gemm_bwd_data")
    ref = as_f64(g_hi, g_lo) @ as_f64(w_hi, w_lo).t()
    tol = dict(rtol=1e-2, atol=1e-2) if dtype_name == "bf16" else F32_TOL   # bf16 output rounding
    torch.testing.assert_close(d_a.double(), ref, **tol)


@pytest.mark.parametrize("dtype_name", ["bf16", "fp32"])
@pytest.mark.parametrize("m,d", [(100, 64), (5000, 128), (20000, 512)])
def test_gemm_bwd_weight(cuda, dtype_name, m, d):
    ffi = _ffi()
    dtype = ffi.PB_BF16 if dtype_name == "bf16" else ffi.PB_F32
    k = 7 * d
    gen = torch.Generator(device="cpu").manual_seed(3 * m + d)
    a = torch.randn(m, k, generator=gen).to(cuda)
    g = (torch.randn(m, d, generator=gen) / np.sqrt(m)).to(cuda)
    a_hi, a_lo = operands(a, dtype)
    g_hi, g_lo = operands(g, dtype)
    ws_bytes = ffi.lib().pb_rgcn_gemm_bwd_weight_workspace_bytes(m, d, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    d_w = torch.full((k, d), float("nan"), device=cuda)
    ffi.check(ffi.lib().pb_rgcn_gemm_bwd_weight(ptr(a_hi), ptr(a_lo), k, ptr(g_hi), ptr(g_lo), d, ptr(d_w), m, d, k,
                                                None, dtype, ptr(ws), ws_bytes, st()), "pb_rgcn_gemm_bwd_weight")
    ref = as_f64(a_hi, a_lo).t() @ as_f64(g_hi, g_lo)
    tol = dict(rtol=2e-3, atol=2e-3) if dtype_name == "bf16" else F32_TOL
    torch.testing.assert_close(d_w.double(), ref, **tol)
    # determinism of the split-K reduction
    d_w2 = torch.empty_like(d_w)
    ffi.check(ffi.lib().pb_rgcn_gemm_bwd_weight(ptr(a_hi), ptr(a_lo), k, ptr(g_hi), ptr(g_lo), d, ptr(d_w2), m, d, k,
                                                None, dtype, ptr(ws), ws_bytes, st()), "pb_rgcn_gemm_bwd_weight")
    assert torch.equal(d_w, d_w2)


def test_weight_prep(cuda):
    ffi = _ffi()
    r, d = 6, 128
    w = torch.randn(r, d, d, device=cuda)
    root = torch.randn(d, d, device=cuda)
    wcat = torch.cat((w.view(r * d, d), root), 0)
    k = (r + 1) * d
    w_bf, wt_bf = torch.empty((k, d), dtype=torch.bfloat16, device=cuda), torch.empty((d, k), dtype=torch.bfloat16, device=cuda)
    ffi.check(ffi.lib().pb_weight_prep(ptr(w), ptr(root), r, d, ffi.PB_BF16, ptr(w_bf), None, ptr(wt_bf), None, st()), "prep")
    assert torch.equal(w_bf, wcat.to(torch.bfloat16)) and torch.equal(wt_bf, wcat.t().to(torch.bfloat16))
    bufs = [torch.empty((k, d), device=cuda), torch.empty((k, d), device=cuda), torch.empty((d, k), device=cuda),
            torch.empty((d, k), device=cuda)]
    ffi.check(ffi.lib().pb_weight_prep(ptr(w), ptr(root), r, d, ffi.PB_F32, *[ptr(b) for b in bufs], st()), "prep")
    hi, lo = tf32_split(wcat)
    assert torch.equal(bufs[0], hi) and torch.equal(bufs[1], lo)
    assert torch.equal(bufs[2], hi.t()) and torch.equal(bufs[3], lo.t())
    assert torch.equal(bufs[0] + bufs[1], wcat)


# ------------------------------------------------------------------------------------------------- aggregation
def _graph(cuda, bsz=6, n_bars=3, p=0.3, seed=1):
    import polyphemus_b200 as pb

    s_np = go.synthetic_structure(bsz, n_bars, p, seed)
    arrays = go.batch_graph(s_np)
    g = pb.graphs_from_tensor(torch.from_numpy(s_np).to(cuda))
    return g, arrays


def _oracle_h(x, arrays, table, keep=None, p=0.0):
    """H[v, r, :] = mean over segment of relu(x[src]*T[dist]) (float64 on CPU)."""
    n, d = x.shape
    ei = torch.from_numpy(arrays.edge_index)
    et, ed = torch.from_numpy(arrays.edge_type), torch.from_numpy(arrays.edge_dist)
    msg = torch.relu(x[ei[0]] * table[ed])
    if keep is not None:
        msg = msg * keep.to(msg.dtype) / (1 - p)
    h = torch.zeros(n * 6, d, dtype=x.dtype).index_add_(0, ei[1] * 6 + et, msg)
    cnt = torch.zeros(n * 6, dtype=x.dtype).index_add_(0, ei[1] * 6 + et, torch.ones(ei.shape[1], dtype=x.dtype))
    return (h / cnt.clamp(min=1).unsqueeze(1)).view(n, 6 * d)


@pytest.mark.parametrize("d", [64, 192, 256, 512, 1024])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_agg_fwd(cuda, d, p_drop):
    ffi = _ffi()
    from polyphemus_b200 import ops

    g, arrays = _graph(cuda)
    n, k = g.num_nodes, 7 * d
    gen = torch.Generator().manual_seed(d)
    x = torch.randn(n, d, generator=gen)
    nn_w, nn_b = torch.randn(d, 32, generator=gen) * 0.3, torch.randn(d, generator=gen) * 0.3
    table = torch.empty(32, d, device=cuda)
    nn_w_dev, nn_b_dev = nn_w.to(cuda), nn_b.to(cuda)          # named: must outlive the asynchronous launch
    ffi.check(ffi.lib().pb_edge_table_fwd(ptr(nn_w_dev), ptr(nn_b_dev), d, ptr(table), st()), "table")
    torch.testing.assert_close(table.cpu(), nn_w.t() + nn_b, rtol=0, atol=0)
    seed = 1234567
    keep = ops.dropout_keep_mask(arrays.edge_index.shape[1], d, p_drop, seed, cuda).cpu() if p_drop else None
    if p_drop:
        assert abs(keep.float().mean().item() - (1 - p_drop)) < 0.01
    ref = torch.cat((_oracle_h(x.double(), arrays, table.cpu().double(), keep, p_drop), x.double()), 1)
    xd = x.to(cuda)
    a_hi, a_lo = torch.empty(n, k, device=cuda), torch.empty(n, k, device=cuda)
    bits = keep_bits(arrays.edge_index.shape[1], d, p_drop, seed, cuda)
    ffi.check(ffi.lib().pb_agg_fwd(g.plan.ref(), ptr(xd), d, ptr(table), ptr(a_hi), ptr(a_lo), k, ffi.PB_F32, ptr(bits), p_drop,
                                   ffi.PB_F32, st()), "agg_fwd")
    torch.testing.assert_close((a_hi.double() + a_lo.double()).cpu(), ref, rtol=2e-6, atol=1e-6)
    assert ((a_hi.view(torch.int32) & 8191) == 0).all()          # hi is a clean TF32 value
    a_bf = torch.empty(n, k, dtype=torch.bfloat16, device=cuda)
    ffi.check(ffi.lib().pb_agg_fwd(g.plan.ref(), ptr(xd), d, ptr(table), ptr(a_bf), None, k, ffi.PB_BF16, ptr(bits), p_drop,
                                   ffi.PB_F32, st()), "agg_fwd")
    torch.testing.assert_close(a_bf.float().cpu(), ref.float(), rtol=8e-3, atol=1e-6)
    # bf16 activation storage: the same kernel reading bf16 node features == the fp32 kernel on the rounded features
    x_bf = xd.to(torch.bfloat16)
    a_bf2, a_bf3 = torch.empty_like(a_bf), torch.empty_like(a_bf)
    ffi.check(ffi.lib().pb_agg_fwd(g.plan.ref(), ptr(x_bf), d, ptr(table), ptr(a_bf2), None, k, ffi.PB_BF16, ptr(bits), p_drop,
                                   ffi.PB_BF16, st()), "agg_fwd")
    x_rounded = x_bf.float()
    ffi.check(ffi.lib().pb_agg_fwd(g.plan.ref(), ptr(x_rounded), d, ptr(table), ptr(a_bf3), None, k, ffi.PB_BF16, ptr(bits),
                                   p_drop, ffi.PB_F32, st()), "agg_fwd")
    assert torch.equal(a_bf2, a_bf3)
    # bit-reproducible
    a2 = torch.empty_like(a_hi)
    l2 = torch.empty_like(a_lo)
    ffi.check(ffi.lib().pb_agg_fwd(g.plan.ref(), ptr(xd), d, ptr(table), ptr(a2), ptr(l2), k, ffi.PB_F32, ptr(bits), p_drop,
                                   ffi.PB_F32, st()), "agg_fwd")
    assert torch.equal(a2, a_hi) and torch.equal(l2, a_lo)


@pytest.mark.parametrize("d", [64, 192, 256, 512, 1024])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_agg_bwd(cuda, d, p_drop):
    ffi = _ffi()
    from polyphemus_b200 import ops

    g, arrays = _graph(cuda, bsz=5, n_bars=2, p=0.35, seed=d)
    n, k = g.num_nodes, 7 * d
    gen = torch.Generator().manual_seed(d + 1)
    x = torch.randn(n, d, generator=gen, dtype=torch.float64).float().double().requires_grad_(True)
    table = (torch.randn(32, d, generator=gen, dtype=torch.float64) * 0.5).float().double().requires_grad_(True)
    d_a = torch.randn(n, k, generator=gen, dtype=torch.float64).float().double()
    gy = torch.randn(n, d, generator=gen, dtype=torch.float64).float().double()
    seed = 42
    keep = ops.dropout_keep_mask(arrays.edge_index.shape[1], d, p_drop, seed, cuda).cpu() if p_drop else None
    a_ref = torch.cat((_oracle_h(x, arrays, table, keep, p_drop), x), 1)
    (a_ref * d_a).sum().backward()
    gx_ref = x.grad + gy
    n_items = g.plan.n_dist_items
    x_dev, t_dev, gy_dev = x.detach().float().to(cuda), table.detach().float().to(cuda), gy.float().to(cuda)
    bits = keep_bits(arrays.edge_index.shape[1], d, p_drop, seed, cuda)
    n_edges = arrays.edge_index.shape[1]
    for dtype, tol in ((ffi.PB_F32, dict(rtol=1e-5, atol=1e-5)), (ffi.PB_BF16, dict(rtol=2e-2, atol=2e-2))):
        da_dev = d_a.float().to(cuda) if dtype == ffi.PB_F32 else d_a.to(torch.bfloat16).to(cuda)
        gx = torch.empty(n, d, device=cuda)
        q_buf = torch.empty(n_edges, d, dtype=da_dev.dtype, device=cuda)
        parts = torch.empty(n_items, d, device=cuda)
        ffi.check(ffi.lib().pb_agg_bwd(g.plan.ref(), ptr(x_dev), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gy_dev), ptr(gx),
                                       ptr(q_buf), ptr(parts), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd")
        torch.testing.assert_close(gx.double().cpu(), gx_ref, **tol)
        if dtype == ffi.PB_BF16:
            # bf16 activation storage == the fp32-storage kernel on the rounded inputs, output rounded to bf16
            xb, gyb = x_dev.to(torch.bfloat16), gy_dev.to(torch.bfloat16)
            gxb, partsb = torch.empty(n, d, dtype=torch.bfloat16, device=cuda), torch.empty_like(parts)
            ffi.check(ffi.lib().pb_agg_bwd(g.plan.ref(), ptr(xb), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gyb), ptr(gxb),
                                           ptr(q_buf), ptr(partsb), ptr(bits), p_drop, ffi.PB_BF16, st()), "agg_bwd")
            xr, gyr = xb.float(), gyb.float()
            gxr, partsr = torch.empty_like(gx), torch.empty_like(parts)
            ffi.check(ffi.lib().pb_agg_bwd(g.plan.ref(), ptr(xr), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gyr), ptr(gxr),
                                           ptr(q_buf), ptr(partsr), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd")
            assert torch.equal(gxb, gxr.to(torch.bfloat16)) and torch.equal(partsb, partsr)
        g_w, g_b = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
        ffi.check(ffi.lib().pb_edge_table_bwd(ptr(parts), ptr(g.plan.dist_item_ptr), d, ptr(g_w), ptr(g_b), st()), "table_bwd")
        scale = float(table.grad.abs().max())
        torch.testing.assert_close(g_w.double().cpu(), table.grad.t(), rtol=tol["rtol"], atol=tol["atol"] * max(1.0, scale))
        torch.testing.assert_close(g_b.double().cpu(), table.grad.sum(0), rtol=tol["rtol"], atol=tol["atol"] * 8 * max(1.0, scale))
        if dtype == ffi.PB_F32:
            gx2, parts2 = torch.empty_like(gx), torch.empty_like(parts)
            ffi.check(ffi.lib().pb_agg_bwd(g.plan.ref(), ptr(x_dev), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gy_dev),
                                           ptr(gx2), ptr(q_buf), ptr(parts2), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd")
            assert torch.equal(gx, gx2) and torch.equal(parts, parts2)
        # ---- the fused kernel (product path): same gx bit for bit (same additions in the same order), the table
        # gradient accumulated without the E x d intermediate
        n_part = ffi.lib().pb_agg_bwd_num_partials(n, d, dtype)
        fparts = torch.full((n_part, 32, d), float("nan"), device=cuda)
        gxf = torch.full((n, d), float("nan"), device=cuda)
        ffi.check(ffi.lib().pb_agg_bwd_fused(g.plan.ref(), ptr(x_dev), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gy_dev), ptr(gxf),
                                             ptr(fparts), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd_fused")
        assert torch.equal(gxf, gx)
        fg_w, fg_b = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
        ffi.check(ffi.lib().pb_edge_table_bwd_fused(ptr(fparts), n_part, d, ptr(fg_w), ptr(fg_b), st()), "table_bwd_fused")
        # fp32 accumulation of unrounded rows: tighter than the legacy path's bf16 intermediate
        ftol = dict(rtol=1e-5, atol=1e-5) if dtype == ffi.PB_F32 else dict(rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(fg_w.double().cpu(), table.grad.t(), rtol=ftol["rtol"], atol=ftol["atol"] * max(1.0, scale))
        torch.testing.assert_close(fg_b.double().cpu(), table.grad.sum(0), rtol=ftol["rtol"], atol=ftol["atol"] * 8 * max(1.0, scale))
        gxf2, fparts2 = torch.empty_like(gxf), torch.empty_like(fparts)
        ffi.check(ffi.lib().pb_agg_bwd_fused(g.plan.ref(), ptr(x_dev), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gy_dev), ptr(gxf2),
                                             ptr(fparts2), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd_fused")
        assert torch.equal(gxf, gxf2) and torch.equal(fparts, fparts2)          # bit-reproducible
        if dtype == ffi.PB_BF16:
            xb, gyb = x_dev.to(torch.bfloat16), gy_dev.to(torch.bfloat16)
            gxb, fpartsb = torch.empty(n, d, dtype=torch.bfloat16, device=cuda), torch.empty_like(fparts)
            ffi.check(ffi.lib().pb_agg_bwd_fused(g.plan.ref(), ptr(xb), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gyb), ptr(gxb),
                                                 ptr(fpartsb), ptr(bits), p_drop, ffi.PB_BF16, st()), "agg_bwd_fused")
            xr, gyr = xb.float(), gyb.float()
            gxr, fpartsr = torch.empty_like(gxf), torch.empty_like(fparts)
            ffi.check(ffi.lib().pb_agg_bwd_fused(g.plan.ref(), ptr(xr), d, ptr(t_dev), ptr(da_dev), k, dtype, ptr(gyr), ptr(gxr),
                                                 ptr(fpartsr), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd_fused")
            assert torch.equal(gxb, gxr.to(torch.bfloat16)) and torch.equal(fpartsb, fpartsr)
        # no residual branch (plain GCL without BatchNorm)
        gxn = torch.empty_like(gxf)
        ffi.check(ffi.lib().pb_agg_bwd_fused(g.plan.ref(), ptr(x_dev), d, ptr(t_dev), ptr(da_dev), k, dtype, None, ptr(gxn),
                                             ptr(fparts2), ptr(bits), p_drop, ffi.PB_F32, st()), "agg_bwd_fused")
        torch.testing.assert_close(gxn.double().cpu(), gx_ref - gy, **tol)


@pytest.mark.parametrize("d", [256, 512])
def test_agg_bwd_tensor_core_large_graph_matches_legacy(cuda, d):
    """The tensor-core backward (bf16 mode, d in {256, 512}) on a graph big enough that every CTA wraps its 4-stage ring
    many times (LMD16, 48 sequences: ~25k nodes, ~85k edges -> ~18 stages per CTA), structured layout with node_order
    and padding rows: gx bit-identical to the legacy two-kernel path, table gradient equal to fp32 accumulation
    order, bit-reproducible."""
    import polyphemus_b200 as pb
    ffi = _ffi()
    s_np = go.synthetic_structure(48, 16, 0.25, seed=9)
    g = pb.graphs_from_tensor(torch.from_numpy(s_np).to(cuda))
    stp = g.structured
    plan, n, k, e = stp.plan, stp.n_padded, 4 * d, g.num_edges
    gen = torch.Generator().manual_seed(d)
    valid = torch.zeros(n, dtype=torch.bool)
    valid[stp.pos.cpu()] = True
    x = (torch.randn(n, d, generator=gen) * valid.unsqueeze(1)).to(cuda).to(torch.bfloat16)
    gy = (torch.randn(n, d, generator=gen) * valid.unsqueeze(1)).to(cuda).to(torch.bfloat16)
    d_a = (torch.randn(n, k, generator=gen) * valid.unsqueeze(1)).to(cuda).to(torch.bfloat16)
    table = (torch.randn(32, d, generator=gen) * 0.5).to(cuda)
    p_drop, seed = 0.1, 77
    bits = keep_bits(e, d, p_drop, seed, cuda)
    lib = ffi.lib()
    # legacy
    gx0 = torch.empty(n, d, dtype=torch.bfloat16, device=cuda)
    q_buf = torch.empty(e, d, dtype=torch.bfloat16, device=cuda)
    parts0 = torch.empty(plan.n_dist_items, d, device=cuda)
    ffi.check(lib.pb_agg_bwd(plan.ref(), ptr(x), d, ptr(table), ptr(d_a), k, ffi.PB_BF16, ptr(gy), ptr(gx0), ptr(q_buf), ptr(parts0),
                             ptr(bits), p_drop, ffi.PB_BF16, st()), "agg_bwd")
    gw0, gb0 = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
    ffi.check(lib.pb_edge_table_bwd(ptr(parts0), ptr(plan.dist_item_ptr), d, ptr(gw0), ptr(gb0), st()), "table_bwd")
    # tensor-core path
    n_part = lib.pb_agg_bwd_num_partials(n, d, ffi.PB_BF16)
    outs = []
    for _ in range(2):
        gx1 = torch.full((n, d), float("nan"), dtype=torch.bfloat16, device=cuda)
        parts1 = torch.full((n_part, 32, d), float("nan"), device=cuda)
        ffi.check(lib.pb_agg_bwd_fused(plan.ref(), ptr(x), d, ptr(table), ptr(d_a), k, ffi.PB_BF16, ptr(gy), ptr(gx1), ptr(parts1),
                                       ptr(bits), p_drop, ffi.PB_BF16, st()), "agg_bwd_fused")
        gw1, gb1 = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
        ffi.check(lib.pb_edge_table_bwd_fused(ptr(parts1), n_part, d, ptr(gw1), ptr(gb1), st()), "table_bwd_fused")
        outs.append((gx1, parts1, gw1, gb1))
    (gx1, parts1, gw1, gb1), (gx2, parts2, gw2, gb2) = outs
    assert torch.equal(gx1, gx0)
    assert torch.equal(gx1, gx2) and torch.equal(parts1, parts2) and torch.equal(gw1, gw2)
    # both paths sum the same bf16-rounded rows in fp32, in different orders
    scale = float(gw0.abs().max())
    torch.testing.assert_close(gw1, gw0, rtol=1e-4, atol=2e-5 * scale)
    torch.testing.assert_close(gb1, gb0, rtol=1e-4, atol=2e-4 * scale)
    # and against the exact sum of the rows the legacy path left in q_buf
    ref = torch.zeros(32, d, dtype=torch.float64, device=cuda)
    dist_of_pos = ((plan.out_rec[:e, 1] >> 8) & 31).long()
    ref.index_add_(0, dist_of_pos, q_buf.double())
    torch.testing.assert_close(gw1.double().t(), ref, rtol=1e-4, atol=2e-5 * scale)


def test_agg_bwd_fused_foreign_graph_high_degree(cuda):
    """Arbitrary edge list (out-degree far above the 8 pipelined slots, self-loops, isolated nodes): the unpipelined
    tail of the fused backward and the generic 6-relation plan against autograd in fp64."""
    from types import SimpleNamespace
    from polyphemus_b200.graph import CsrPlan

    ffi = _ffi()
    n, e, d = 150, 4000, 128
    gen = torch.Generator().manual_seed(5)
    ei = torch.randint(0, n - 10, (2, e), generator=gen)            # the last 10 nodes are isolated
    ei[:, :50] = 3                                                  # 50 parallel self-loops on node 3
    et = torch.randint(0, 6, (e,), generator=gen, dtype=torch.uint8)
    ed = torch.randint(0, 32, (e,), generator=gen, dtype=torch.uint8)
    arrays = SimpleNamespace(edge_index=ei.numpy(), edge_type=et.long().numpy(), edge_dist=ed.long().numpy())
    plan = CsrPlan(ei.to(cuda), et.to(cuda), ed.to(cuda), n)
    k = 7 * d
    x = torch.randn(n, d, generator=gen, dtype=torch.float64).float().double().requires_grad_(True)
    table = (torch.randn(32, d, generator=gen, dtype=torch.float64) * 0.5).float().double().requires_grad_(True)
    d_a = torch.randn(n, k, generator=gen, dtype=torch.float64).float().double()
    gy = torch.randn(n, d, generator=gen, dtype=torch.float64).float().double()
    a_ref = torch.cat((_oracle_h(x, arrays, table), x), 1)
    (a_ref * d_a).sum().backward()
    x_dev, t_dev, gy_dev, da_dev = (v.detach().float().to(cuda) for v in (x, table, gy, d_a))
    n_part = ffi.lib().pb_agg_bwd_num_partials(n, d, ffi.PB_F32)
    parts = torch.empty((n_part, 32, d), device=cuda)
    gx = torch.empty(n, d, device=cuda)
    ffi.check(ffi.lib().pb_agg_bwd_fused(plan.ref(), ptr(x_dev), d, ptr(t_dev), ptr(da_dev), k, ffi.PB_F32, ptr(gy_dev), ptr(gx),
                                         ptr(parts), None, 0.0, ffi.PB_F32, st()), "agg_bwd_fused")
    torch.testing.assert_close(gx.double().cpu(), x.grad + gy, rtol=1e-5, atol=1e-5)
    g_w, g_b = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
    ffi.check(ffi.lib().pb_edge_table_bwd_fused(ptr(parts), n_part, d, ptr(g_w), ptr(g_b), st()), "table_bwd_fused")
    scale = max(1.0, float(table.grad.abs().max()))
    torch.testing.assert_close(g_w.double().cpu(), table.grad.t(), rtol=1e-5, atol=1e-5 * scale)


@pytest.mark.parametrize("act", ["bf16", "fp32"])
@pytest.mark.parametrize("d", [256, 512])
def test_agg_bwd_tensor_core_foreign_graph_high_degree(cuda, act, d):
    """Both tensor-core backward kernels (bf16 rows: cp.async ring kernel; fp32 rows: register-gather kernel) on an
    arbitrary edge list with the generic 6-relation plan: a hub with 400 out-edges (more than the 128 ring slots: the
    synchronous 32-edge chunks wrap the stages inside ONE source), 50 parallel self-loops, isolated nodes, dropout on.
    gx bit-identical to the legacy two-kernel path, table gradient equal up to fp32 summation order."""
    from polyphemus_b200.graph import CsrPlan

    ffi = _ffi()
    lib = ffi.lib()
    n, e = 300, 6400
    gen = torch.Generator().manual_seed(11 + d)
    ei = torch.randint(0, n - 10, (2, e), generator=gen)            # the last 10 nodes are isolated
    ei[:, :50] = 3                                                  # 50 parallel self-loops on node 3
    ei[0, 50:450] = 5                                               # hub: 400 out-edges of node 5
    et = torch.randint(0, 6, (e,), generator=gen, dtype=torch.uint8)
    ed = torch.randint(0, 32, (e,), generator=gen, dtype=torch.uint8)
    plan = CsrPlan(ei.to(cuda), et.to(cuda), ed.to(cuda), n)
    k = 7 * d
    adt = torch.bfloat16 if act == "bf16" else torch.float32
    acode = ffi.PB_BF16 if act == "bf16" else ffi.PB_F32
    x = torch.randn(n, d, generator=gen).to(cuda).to(adt)
    gy = torch.randn(n, d, generator=gen).to(cuda).to(adt)
    d_a = torch.randn(n, k, generator=gen).to(cuda).to(torch.bfloat16)
    table = (torch.randn(32, d, generator=gen) * 0.5).to(cuda)
    p_drop = 0.1
    bits = keep_bits(e, d, p_drop, 4242, cuda)
    gx0 = torch.empty(n, d, dtype=adt, device=cuda)
    q_buf = torch.empty(e, d, dtype=torch.bfloat16, device=cuda)
    parts0 = torch.empty(plan.n_dist_items, d, device=cuda)
    ffi.check(lib.pb_agg_bwd(plan.ref(), ptr(x), d, ptr(table), ptr(d_a), k, ffi.PB_BF16, ptr(gy), ptr(gx0), ptr(q_buf), ptr(parts0),
                             ptr(bits), p_drop, acode, st()), "agg_bwd")
    gw0, gb0 = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
    ffi.check(lib.pb_edge_table_bwd(ptr(parts0), ptr(plan.dist_item_ptr), d, ptr(gw0), ptr(gb0), st()), "table_bwd")
    n_part = lib.pb_agg_bwd_num_partials(n, d, ffi.PB_BF16)
    outs = []
    for _ in range(2):
        gx1 = torch.full((n, d), float("nan"), dtype=adt, device=cuda)
        parts1 = torch.full((n_part, 32, d), float("nan"), device=cuda)
        ffi.check(lib.pb_agg_bwd_fused(plan.ref(), ptr(x), d, ptr(table), ptr(d_a), k, ffi.PB_BF16, ptr(gy), ptr(gx1), ptr(parts1),
                                       ptr(bits), p_drop, acode, st()), "agg_bwd_fused")
        gw1, gb1 = torch.empty(d, 32, device=cuda), torch.empty(d, device=cuda)
        ffi.check(lib.pb_edge_table_bwd_fused(ptr(parts1), n_part, d, ptr(gw1), ptr(gb1), st()), "table_bwd_fused")
        outs.append((gx1, parts1, gw1))
    (gx1, parts1, gw1), (gx2, parts2, gw2) = outs
    assert torch.equal(gx1, gx0)
    assert torch.equal(gx1, gx2) and torch.equal(parts1, parts2) and torch.equal(gw1, gw2)
    scale = float(gw0.abs().max())
    torch.testing.assert_close(gw1, gw0, rtol=1e-4, atol=2e-5 * scale)
    if act == "bf16":
        # forward on the same graph (in-degree > 32 on node 3): the pipelined kernel (d = 512, bf16 rows) and the generic one
        # (same values as fp32 rows) compute the same operand up to the rounding of the dropout scaling and the bf16 store
        a_b = torch.full((n, k), float("nan"), dtype=torch.bfloat16, device=cuda)
        a_f = torch.full((n, k), float("nan"), dtype=torch.bfloat16, device=cuda)
        ffi.check(lib.pb_agg_fwd(plan.ref(), ptr(x), d, ptr(table), ptr(a_b), None, k, ffi.PB_BF16, ptr(bits), p_drop, ffi.PB_BF16, st()), "agg_fwd")
        x32 = x.float()
        ffi.check(lib.pb_agg_fwd(plan.ref(), ptr(x32), d, ptr(table), ptr(a_f), None, k, ffi.PB_BF16, ptr(bits), p_drop, ffi.PB_F32, st()), "agg_fwd")
        torch.testing.assert_close(a_b.float(), a_f.float(), rtol=1.6e-2, atol=1e-3)


def test_split_rows_linear_matches_two_linears(cuda):
    """ops.split_rows_linear (the content decoder's un-embedding by instrument class): values and all gradients equal two
    F.linear calls on the row blocks, including an empty first / second block."""
    from polyphemus_b200 import ops

    gen = torch.Generator().manual_seed(3)
    m, k, n = 300, 128, 192
    for n0 in (0, 77, 300):
        x = torch.randn(m, k, generator=gen).to(cuda).requires_grad_(True)
        w0, w1 = (torch.randn(n, k, generator=gen).to(cuda).requires_grad_(True) for _ in range(2))
        b0, b1 = (torch.randn(n, generator=gen).to(cuda).requires_grad_(True) for _ in range(2))
        o0, o1 = ops.split_rows_linear(x, n0, w0, b0, w1, b1, precision="fp32")
        g0, g1 = torch.randn(n0, n, generator=gen).to(cuda), torch.randn(m - n0, n, generator=gen).to(cuda)
        ((o0 * g0).sum() + (o1 * g1).sum()).backward()
        ref = [t.detach().double().requires_grad_(True) for t in (x, w0, b0, w1, b1)]
        r0 = torch.nn.functional.linear(ref[0][:n0], ref[1], ref[2])
        r1 = torch.nn.functional.linear(ref[0][n0:], ref[3], ref[4])
        ((r0 * g0.double()).sum() + (r1 * g1.double()).sum()).backward()
        tol = lambda t: dict(rtol=1e-4, atol=1e-5 * max(1.0, float(t.abs().max()) if t.numel() else 1.0))
        torch.testing.assert_close(o0.double(), r0.detach(), **tol(r0))
        torch.testing.assert_close(o1.double(), r1.detach(), **tol(r1))
        for got, want in zip((x, w0, b0, w1, b1), ref):
            want_g = want.grad if want.grad is not None else torch.zeros_like(want)
            torch.testing.assert_close(got.grad.double(), want_g, **tol(want_g))


def test_keep_bits_prefetch_is_bit_identical(cuda):
    """Drawing the GCL dropout keep-bits of a stack ahead on the side stream (ops.prefetch_keep_bits) uses the same seeds in
    the same order as drawing them inside each layer call: outputs and every gradient are bit-identical."""
    import itertools
    import polyphemus_b200 as pb
    from polyphemus_b200 import ops

    g, _ = _graph(cuda, bsz=6, n_bars=4, p=0.3, seed=21)
    pb.set_precision("bf16")
    try:
        results = []
        for prefetch in (True, False):
            torch.manual_seed(5)
            gcn = pb.GCN(input_dim=256, hidden_dim=256, n_layers=3, num_relations=6, batch_norm=True, dropout=0.0).to(cuda).train()
            for layer in gcn.layers:
                layer.dropout = 0.1
            ops._bits_prefetch = prefetch
            ops._seed_counter = itertools.count()
            g.x = torch.randn(g.num_nodes, 256, generator=torch.Generator().manual_seed(1)).to(cuda).requires_grad_(True)
            y = gcn(g)
            y.square().sum().backward()
            results.append([y.detach().clone(), g.x.grad.clone()] + [p.grad.clone() for p in gcn.parameters()])
        for a, b in zip(*results):
            assert torch.equal(a, b)
    finally:
        ops._bits_prefetch = True
        pb.set_precision("fp32")


# ------------------------------------------------------------------------------------------------- BatchNorm
@pytest.mark.parametrize("m,d", [(37, 64), (4000, 512), (70001, 256)])
def test_bn_relu_res_fwd_bwd(cuda, m, d):
    ffi = _ffi()
    lib = ffi.lib()
    gen = torch.Generator().manual_seed(m)
    out = (torch.randn(m, d, generator=gen) * 2 + 50.0)         # large mean: the variance must not cancel
    x = torch.randn(m, d, generator=gen)
    gamma, beta = torch.rand(d, generator=gen) + 0.5, torch.randn(d, generator=gen)
    rm, rv = torch.randn(d, generator=gen), torch.rand(d, generator=gen) + 0.5
    gy = torch.randn(m, d, generator=gen)
    # fp64 reference of the forward through torch (training-mode batch statistics)
    rm64, rv64 = rm.double().clone(), rv.double().clone()
    y_ref = x.double() + torch.relu(torch.nn.functional.batch_norm(out.double(), rm64, rv64, gamma.double(), beta.double(),
                                                                   True, 0.1, 1e-5))
    dev = lambda t: t.to(cuda).contiguous()
    out_d, x_d, gamma_d, beta_d, rm_d, rv_d, gy_d = map(dev, (out, x, gamma, beta, rm, rv, gy))
    ws_bytes = lib.pb_bn_workspace_bytes(m, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    coef, save = torch.empty(3, d, device=cuda), torch.empty(2, d, device=cuda)
    ffi.check(lib.pb_bn_stats(ptr(out_d), d, m, d, None, ptr(gamma_d), ptr(beta_d), 1e-5, 0.1, ptr(rm_d), ptr(rv_d),
                              ptr(save), ptr(coef), ptr(ws), ws_bytes, ffi.PB_F32, st()), "bn_stats")
    y = torch.empty(m, d, device=cuda)
    ffi.check(lib.pb_bn_relu_res_fwd(ptr(out_d), d, ptr(x_d), ptr(coef), ptr(y), m, d, None, 1, ffi.PB_F32, st()), "bn_fwd")
    torch.testing.assert_close(y.double().cpu(), y_ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(rm_d.double().cpu(), rm64, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rv_d.double().cpu(), rv64, rtol=1e-4, atol=1e-5)
    mean64, var64 = out.double().mean(0), out.double().var(0, unbiased=False)
    torch.testing.assert_close(save[0].double().cpu(), mean64, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(save[1].double().cpu(), 1 / torch.sqrt(var64 + 1e-5), rtol=1e-5, atol=0)
    # fp64 reference of the backward. The ReLU mask is a step function: an element whose pre-activation is
    # within rounding of 0 may legitimately flip and then moves its whole column's reductions, so the reference
    # takes the mask exactly as the kernel evaluates it: fma(fl32(out - mean), scale, beta) > 0.
    mu32, sc32, be32 = coef[0].cpu(), coef[1].cpu(), coef[2].cpu()
    ctr32 = out - mu32
    mask = (ctr32.double() * sc32.double() + be32.double()) > 0
    xhat = ctr32.double() * save[1].double().cpu()
    gz = gy.double() * mask
    g_beta_ref, g_gamma_ref = gz.sum(0), (gz * xhat).sum(0)
    g_out_ref = sc32.double() * (gz - g_beta_ref / m - xhat * g_gamma_ref / m)
    g_hi, g_lo = torch.empty(m, d, device=cuda), torch.empty(m, d, device=cuda)
    g_gamma, g_beta, g_bias = (torch.empty(d, device=cuda) for _ in range(3))
    ffi.check(lib.pb_bn_relu_res_bwd(ptr(gy_d), ptr(out_d), d, ptr(gamma_d), ptr(save), ptr(coef), m, d, None, ffi.PB_F32,
                                     ptr(g_hi), ptr(g_lo), d, ptr(g_gamma), ptr(g_beta), ptr(g_bias), ptr(ws), ws_bytes,
                                     ffi.PB_F32, st()), "bn_bwd")
    scale = float(g_out_ref.abs().max())
    torch.testing.assert_close((g_hi.double() + g_lo.double()).cpu(), g_out_ref, rtol=1e-4, atol=1e-5 * max(1.0, scale))
    torch.testing.assert_close(g_gamma.double().cpu(), g_gamma_ref, rtol=1e-4, atol=1e-4 * float(g_gamma_ref.abs().max()))
    torch.testing.assert_close(g_beta.double().cpu(), g_beta_ref, rtol=1e-4, atol=1e-4 * float(g_beta_ref.abs().max()))
    assert g_bias.abs().max().item() < 1e-3 * max(1.0, scale) * np.sqrt(m)      # column sums of a BN gradient vanish
    # and the formula itself against autograd on a kink-free subset of columns
    o64 = out.double().requires_grad_(True)
    g64 = gamma.double().requires_grad_(True)
    z64 = torch.nn.functional.batch_norm(o64, None, None, g64, beta.double(), True, 0.1, 1e-5)
    (x.double() + torch.relu(z64)).backward(gy.double())
    cols = (z64.detach().abs() > 1e-3).all(dim=0)
    if cols.any():
        torch.testing.assert_close(g_out_ref[:, cols], o64.grad[:, cols], rtol=1e-4, atol=1e-5 * max(1.0, scale))
    g_bf = torch.empty(m, d, dtype=torch.bfloat16, device=cuda)
    ffi.check(lib.pb_bn_relu_res_bwd(ptr(gy_d), ptr(out_d), d, ptr(gamma_d), ptr(save), ptr(coef), m, d, None, ffi.PB_BF16,
                                     ptr(g_bf), None, d, ptr(g_gamma), ptr(g_beta), ptr(g_bias), ptr(ws), ws_bytes,
                                     ffi.PB_F32, st()), "bn_bwd")
    torch.testing.assert_close(g_bf.double().cpu(), g_out_ref, rtol=1e-2, atol=1e-2 * max(1.0, scale))
    # bf16 activation storage: statistics / forward / backward on bf16 `out`, `x`, `gy` == the fp32-storage kernels on
    # the same (rounded) values; y comes back rounded to bf16
    out_b, x_b, gy_b = out_d.to(torch.bfloat16), x_d.to(torch.bfloat16), gy_d.to(torch.bfloat16)
    out_r, x_r, gy_r = out_b.float(), x_b.float(), gy_b.float()
    res = []
    for o_t, x_t, g_t, act in ((out_b, x_b, gy_b, ffi.PB_BF16), (out_r, x_r, gy_r, ffi.PB_F32)):
        rm2, rv2 = dev(rm), dev(rv)
        coef2, save2 = torch.empty(3, d, device=cuda), torch.empty(2, d, device=cuda)
        ffi.check(lib.pb_bn_stats(ptr(o_t), d, m, d, None, ptr(gamma_d), ptr(beta_d), 1e-5, 0.1, ptr(rm2), ptr(rv2),
                                  ptr(save2), ptr(coef2), ptr(ws), ws_bytes, act, st()), "bn_stats")
        y2 = torch.empty(m, d, dtype=o_t.dtype, device=cuda)
        ffi.check(lib.pb_bn_relu_res_fwd(ptr(o_t), d, ptr(x_t), ptr(coef2), ptr(y2), m, d, None, 1, act, st()), "bn_fwd")
        g2 = torch.empty(m, d, dtype=torch.bfloat16, device=cuda)
        gg2, gb2, gbias2 = (torch.empty(d, device=cuda) for _ in range(3))
        ffi.check(lib.pb_bn_relu_res_bwd(ptr(g_t), ptr(o_t), d, ptr(gamma_d), ptr(save2), ptr(coef2), m, d, None, ffi.PB_BF16,
                                         ptr(g2), None, d, ptr(gg2), ptr(gb2), ptr(gbias2), ptr(ws), ws_bytes, act, st()),
                  "bn_bwd")
        res.append((coef2, save2, rm2, rv2, y2.to(torch.bfloat16), g2, gg2, gb2))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # eval mode coefficients
    coef_e = torch.empty(3, d, device=cuda)
    ffi.check(lib.pb_bn_prepare_eval(ptr(gamma_d), ptr(beta_d), ptr(rm_d), ptr(rv_d), 1e-5, d, ptr(coef_e), st()), "bn_eval")
    ffi.check(lib.pb_bn_relu_res_fwd(ptr(out_d), d, ptr(x_d), ptr(coef_e), ptr(y), m, d, None, 1, ffi.PB_F32, st()), "bn_fwd")
    y_eval = x.double() + torch.relu(torch.nn.functional.batch_norm(out.double(), rm_d.double().cpu(), rv_d.double().cpu(),
                                                                    gamma.double(), beta.double(), False, 0.1, 1e-5))
    torch.testing.assert_close(y.double().cpu(), y_eval, rtol=1e-4, atol=2e-5)


def test_grad_prep(cuda):
    ffi = _ffi()
    m, d = 3000, 128
    g = torch.randn(m, d, device=cuda)
    ws_bytes = ffi.lib().pb_bn_workspace_bytes(m, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    hi, lo, gb = torch.empty_like(g), torch.empty_like(g), torch.empty(d, device=cuda)
    ffi.check(ffi.lib().pb_grad_prep(ptr(g), d, m, d, ffi.PB_F32, ptr(hi), ptr(lo), d, ptr(gb), ptr(ws), ws_bytes, st()), "prep")
    assert torch.equal(hi + lo, g)
    torch.testing.assert_close(gb.double(), g.double().sum(0), rtol=1e-5, atol=1e-4)
    bf = torch.empty(m, d, dtype=torch.bfloat16, device=cuda)
    ffi.check(ffi.lib().pb_grad_prep(ptr(g), d, m, d, ffi.PB_BF16, ptr(bf), None, d, ptr(gb), ptr(ws), ws_bytes, st()), "prep")
    assert torch.equal(bf, g.to(torch.bfloat16))


def test_dropout_mask_matches_host_definition(cuda):
    """pb_dropout_mask == the documented counter hash: SplitMix64 mix of seed + GOLDEN*(1 + (eid<<32 | chunk)),
    four 16-bit lanes per 4-channel chunk, keep iff lane >= round(p * 65536)."""
    from polyphemus_b200 import ops

    M = (1 << 64) - 1

    def mix(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    n_edges, d, p, seed = 37, 64, 0.1, 0xDEADBEEF12345
    got = ops.dropout_keep_mask(n_edges, d, p, seed, cuda).cpu().numpy()
    thresh = int(p * 65536.0 + 0.5)
    want = np.zeros((n_edges, d), dtype=bool)
    for e in range(n_edges):
        for c in range(d // 4):
            r = mix((seed + 0x9E3779B97F4A7C15 * (((e << 32) | c) + 1)) & M)
            for i in range(4):
                want[e, 4 * c + i] = ((r >> (16 * i)) & 0xFFFF) >= thresh
    np.testing.assert_array_equal(got, want)
    big = ops.dropout_keep_mask(20000, 512, p, seed, cuda)
    assert abs(float(big.float().mean()) - 0.9) < 2e-3
    # neighbouring channels / edges are uncorrelated
    b = big.float() - big.float().mean()
    assert abs(float((b[:, :-1] * b[:, 1:]).mean())) < 1e-3 and abs(float((b[:-1] * b[1:]).mean())) < 1e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("m,k,n", [(300, 512, 3840), (1000, 7680, 512), (70, 64, 64)])
def test_tensor_core_linear(cuda, precision, m, k, n):
    """ops.tc_linear (chord encoder / decoder shapes, model.py:322,525) against fp64 F.linear, forward + gradients."""
    from polyphemus_b200 import ops

    gen = torch.Generator().manual_seed(m + k + n)
    x = torch.randn(m, k, generator=gen).to(cuda).requires_grad_(True)
    w = (torch.randn(n, k, generator=gen) / np.sqrt(k)).to(cuda).requires_grad_(True)
    b = torch.randn(n, generator=gen).to(cuda).requires_grad_(True)
    gy = torch.randn(m, n, generator=gen).to(cuda)
    y = ops.tc_linear(x, w, b, precision=precision)
    y.backward(gy)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y64 = torch.nn.functional.linear(x64, w64, b64)
    y64.backward(gy.double())
    if precision == "fp32":
        tol = lambda ref: dict(rtol=1e-4, atol=1e-5 * max(1.0, float(ref.abs().max())))
    else:
        tol = lambda ref: dict(rtol=3e-2, atol=3e-2 * float(ref.abs().max()))
    torch.testing.assert_close(y.double(), y64, **tol(y64))
    torch.testing.assert_close(x.grad.double(), x64.grad, **tol(x64.grad))
    torch.testing.assert_close(w.grad.double(), w64.grad, **tol(w64.grad))
    torch.testing.assert_close(b.grad.double(), b64.grad, **tol(b64.grad))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_table_gather_gradient(cuda, precision):
    """ops.table_gather: gather from a tiny table, gradient through the split-K GEMM == index_add reference."""
    from polyphemus_b200 import ops

    gen = torch.Generator().manual_seed(5)
    vocab, c, rows = 131, 128, 40000
    table = torch.randn(vocab, c, generator=gen).to(cuda).requires_grad_(True)
    ids = torch.randint(0, vocab, (rows // 8, 8), generator=gen).to(cuda)
    ids[::3] = 130                                     # a very popular token (PAD)
    g = torch.randn(rows // 8, 8, c, generator=gen).to(cuda)
    out = ops.table_gather(table, ids, precision=precision)
    assert torch.equal(out, table.detach()[ids])
    out.backward(g)
    ref = torch.zeros(vocab, c, dtype=torch.float64, device=cuda).index_add_(0, ids.reshape(-1), g.reshape(-1, c).double())
    if precision == "fp32":
        torch.testing.assert_close(table.grad.double(), ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))
    else:
        torch.testing.assert_close(table.grad.double(), ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("classes,valid", [(192, 131), (128, 99), (64, 64), (320, 300), (512, 512)])
def test_token_nll_matches_cross_entropy(cuda, dtype, classes, valid):
    """ops.token_nll (pb_ce_fwd / pb_ce_bwd) == F.cross_entropy(reduction='none', ignore_index) in fp64 on the same
    (type-rounded) logits, -inf padding columns included; the gradient in the logits' own type."""
    from polyphemus_b200 import ops

    gen = torch.Generator().manual_seed(classes)
    rows, ignore = 10007, valid - 1
    logits = (4.0 * torch.randn(rows, classes, generator=gen)).to(dtype)
    logits[:, valid:] = float("-inf")
    target = torch.randint(0, valid, (rows,), generator=gen)
    target[::5] = ignore
    row_grad = torch.randn(rows, generator=gen)
    lg = logits.to(cuda).requires_grad_(True)
    nll = ops.token_nll(lg, target.int().to(cuda), ignore)
    nll.backward(row_grad.to(cuda))
    ref_in = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, target, ignore_index=ignore, reduction="none")
    ref.backward(row_grad.double())
    torch.testing.assert_close(nll.cpu().double(), ref.detach(), rtol=1e-5, atol=1e-5)
    assert lg.grad.dtype == dtype and torch.isfinite(lg.grad).all()
    assert (lg.grad[::5] == 0).all() and (lg.grad[:, valid:] == 0).all()
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=8e-3, atol=1e-6)
    torch.testing.assert_close(lg.grad.cpu().double(), ref_in.grad, **tol)


def test_token_nll_rejects_bad_layout(cuda):
    from polyphemus_b200 import PolyphemusB200Error, ops

    with pytest.raises(PolyphemusB200Error):
        ops.token_nll(torch.zeros(8, 131, device=cuda), torch.zeros(8, dtype=torch.int32, device=cuda), 130)
    with pytest.raises(ValueError):
        ops.token_nll(torch.zeros(8, 136, device=cuda), torch.zeros(8, dtype=torch.int64, device=cuda), 130)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("d", [64, 192, 512])
def test_chord_embed_matches_dense_formulation(cuda, precision, d):
    """ops.chord_embed (pb_chord_embed_fwd / bwd_prep + split-K GEMM) == relu(bias + sum of table rows) written with
    plain index arithmetic in fp64, and its autograd gradients for the tables and the bias."""
    from polyphemus_b200 import ops
    from polyphemus_b200.train import synthetic_tokens

    gen = torch.Generator().manual_seed(d)
    n, slots, vocab, dur_off = 3001, 15, 230, 131
    tokens = synthetic_tokens(n, gen)                                  # int16 [n, 16, 2]
    is_drum = torch.rand(n, generator=gen) < 0.3
    tables = (0.2 * torch.randn(2, slots, vocab, d, generator=gen)).requires_grad_(True)
    bias = (0.1 * torch.randn(d, generator=gen)).requires_grad_(True)
    g = torch.randn(n, d, generator=gen)

    t_dev = tables.detach().to(cuda).requires_grad_(True)
    b_dev = bias.detach().to(cuda).requires_grad_(True)
    out = ops.chord_embed(t_dev, b_dev, tokens.to(cuda), is_drum.to(cuda), tok_offset=2, dur_off=dur_off, precision=precision)
    out.backward(g.to(cuda))

    tab = tables.detach()
    if precision == "bf16":
        tab = tab.to(torch.bfloat16).float()                           # the forward gathers bf16 tables
    tab = tab.double().requires_grad_(True)
    bias64 = bias.detach().double().requires_grad_(True)
    ids = tokens[:, 1:, :].long()
    sets = is_drum.long().view(-1, 1).expand(-1, slots)
    slot = torch.arange(slots).view(1, -1).expand(n, -1)
    pre = bias64 + tab[sets, slot, ids[..., 0]].sum(1) + tab[sets, slot, dur_off + ids[..., 1]].sum(1)
    ref = torch.relu(pre)
    ref.backward(g.double())
    torch.testing.assert_close(out.detach().cpu().double(), ref.detach(), rtol=1e-5, atol=1e-5)
    scale = float(tab.grad.abs().max())
    tol = dict(rtol=1e-4, atol=1e-5 * scale) if precision == "fp32" else dict(rtol=2e-2, atol=2e-2 * scale)
    torch.testing.assert_close(t_dev.grad.cpu().double(), tab.grad, **tol)
    # fp32 mode: exact fp32 column sums. bf16 mode: the bias gradient is summed from the same bf16-rounded, ReLU-masked
    # operand the table-gradient GEMM contracts (2^-9 relative rounding per term, fp32 accumulation)
    b_tol = 1e-4 if precision == "fp32" else 5e-3
    torch.testing.assert_close(b_dev.grad.cpu().double(), bias64.grad, rtol=b_tol, atol=b_tol * float(bias64.grad.abs().max()))


def _random_bar_ptr(n_bars, gen, max_nodes=128):
    sizes = torch.randint(1, max_nodes + 1, (n_bars,), generator=gen)
    sizes[0], sizes[1] = 1, max_nodes                                    # smallest and largest possible bar
    return torch.cat((torch.zeros(1, dtype=torch.long), sizes.cumsum(0))).int(), sizes


@pytest.mark.parametrize("d", [64, 192, 512, 1024])
def test_bar_pool_matches_global_attention(cuda, d):
    """ops.bar_pool == PyG GlobalAttention (softmax over the bar, weighted sum), forward and both gradients (fp64)."""
    from polyphemus_b200 import ops

    gen = torch.Generator().manual_seed(d)
    bar_ptr, sizes = _random_bar_ptr(37, gen)
    n = int(bar_ptr[-1])
    h = torch.randn(n, d, generator=gen)
    gate = 2.0 * torch.randn(n, generator=gen)
    g_out = torch.randn(37, d, generator=gen)
    seg = torch.repeat_interleave(torch.arange(37), sizes)

    h_dev, gate_dev = h.to(cuda).requires_grad_(True), gate.to(cuda).requires_grad_(True)
    out = ops.bar_pool(h_dev, gate_dev.view(-1, 1), bar_ptr.to(cuda))
    out.backward(g_out.to(cuda))

    h64, gate64 = h.double().requires_grad_(True), gate.double().requires_grad_(True)
    top = torch.full((37,), float("-inf"), dtype=torch.float64).scatter_reduce(0, seg, gate64.detach(), "amax")
    e = torch.exp(gate64 - top[seg])
    alpha = e / (torch.zeros(37, dtype=torch.float64).index_add_(0, seg, e)[seg] + 1e-16)
    ref = torch.zeros(37, d, dtype=torch.float64).index_add_(0, seg, alpha.unsqueeze(1) * h64)
    ref.backward(g_out.double())
    torch.testing.assert_close(out.detach().cpu().double(), ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h_dev.grad.cpu().double(), h64.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gate_dev.grad.cpu().double(), gate64.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("d", [64, 512])
def test_bar_expand_and_segment_sum(cuda, d):
    from polyphemus_b200 import ops

    gen = torch.Generator().manual_seed(d + 1)
    bar_ptr, sizes = _random_bar_ptr(29, gen)
    n = int(bar_ptr[-1])
    seg = torch.repeat_interleave(torch.arange(29), sizes)
    z = torch.randn(29, d, generator=gen)
    g_x = torch.randn(n, d, generator=gen)
    z_dev = z.to(cuda).requires_grad_(True)
    x = ops.bar_expand(z_dev, bar_ptr.to(cuda), n)
    assert torch.equal(x.detach().cpu(), z[seg])
    x.backward(g_x.to(cuda))
    ref = torch.zeros(29, d, dtype=torch.float64).index_add_(0, seg, g_x.double())
    torch.testing.assert_close(z_dev.grad.cpu().double(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_row_scatter_gather_gradients(cuda, out_dtype):
    """ops.ScatterRowsFn / GatherRowsFn (pb_rows_scatter / pb_rows_gather: node -> padded row of the structured layout,
    injective, storage conversion fused): values and gradients equal autograd's index_copy / index_select formulation
    on a real structured plan, padding rows zero."""
    from polyphemus_b200 import ops

    g, _ = _graph(cuda, bsz=5, n_bars=2, p=0.35, seed=7)
    st_ = g.structured
    n, n_rows, d = g.num_nodes, st_.n_padded, 64
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(n, d, generator=gen).to(cuda).requires_grad_(True)
    x_ref = x.detach().clone().requires_grad_(True)
    g_out = torch.randn(n, d, generator=gen).to(cuda)
    w = torch.randn(n_rows, d, generator=gen).to(cuda)
    xp = ops.ScatterRowsFn.apply(x, st_, out_dtype)
    y = ops.GatherRowsFn.apply(xp * w.to(out_dtype), st_, torch.float32)
    y.backward(g_out)
    xp_ref = torch.zeros(n_rows, d, device=cuda).index_copy(0, st_.pos, x_ref).to(out_dtype)
    y_ref = (xp_ref * w.to(out_dtype)).index_select(0, st_.pos).float()
    y_ref.backward(g_out)
    assert xp.dtype == out_dtype and torch.equal(xp, xp_ref) and torch.equal(y, y_ref)
    torch.testing.assert_close(x.grad, x_ref.grad, rtol=0, atol=0)


def test_token_hist_matches_bincount(cuda):
    """pb_token_hist == per-set bincount of the (pitch, duration) ids of slots 1..15 (the histogram that weights the
    BatchNorm statistics of the folded embedding tables, model.py:355-376), incl. a batch without drum nodes."""
    from polyphemus_b200.train import synthetic_tokens
    ffi = _ffi()
    for n, drum_frac in ((5000, 0.3), (777, 0.0), (1, 1.0)):
        tokens = synthetic_tokens(n, torch.Generator().manual_seed(n)).to(cuda)
        is_drum = (torch.rand(n, generator=torch.Generator().manual_seed(n + 1)) < drum_frac).to(cuda)
        counts = torch.full((2, 230), -1, dtype=torch.int64, device=cuda)
        ffi.check(ffi.lib().pb_token_hist(ptr(tokens), 32, 2, 15, ptr(is_drum.view(torch.uint8)), n, 131, 99, ptr(counts), st()), "hist")
        ids = tokens[:, 1:, :].long().cpu()
        for s_ in (0, 1):
            sel = ids[is_drum.cpu() == bool(s_)]
            want = torch.cat((torch.bincount(sel[..., 0].reshape(-1), minlength=131), torch.bincount(sel[..., 1].reshape(-1), minlength=99)))
            assert torch.equal(counts[s_].cpu(), want)


@pytest.mark.parametrize("dtype_name", ["bf16", "fp32"])
@pytest.mark.parametrize("structured", [True, False])
def test_gemm_fwd_bn_partials_and_finalize(cuda, dtype_name, structured):
    """BatchNorm statistics as a by-product of the forward GEMM's epilogue (pb_rgcn_gemm_fwd_bn + pb_bn_finalize) ==
    the separate pass (pb_bn_stats) over the stored output: mean / rstd / running statistics, padding rows excluded."""
    import ctypes
    from test_parity_scale_gpu import _groups
    ffi = _ffi()
    dtype = ffi.PB_BF16 if dtype_name == "bf16" else ffi.PB_F32
    act = dtype                                             # bf16 mode stores `out` in bf16
    d = 512
    gen = torch.Generator().manual_seed(11)
    if structured:
        counts = (300, 0, 129, 1000)
        gs, starts, m = _groups(ffi, counts)
        gref, k = ctypes.byref(gs), 4 * d
        valid = torch.zeros(m, dtype=torch.bool)
        for g in range(4):
            valid[starts[g]:starts[g] + counts[g]] = True
        kw = 7 * d
    else:
        m, k, gref, kw = 1000, 7 * d, None, 7 * d
        valid = torch.ones(m, dtype=torch.bool)
    a = torch.randn(m, k, generator=gen) * valid.unsqueeze(1)
    wt = torch.randn(d, kw, generator=gen) / np.sqrt(k)
    bias = torch.randn(d, generator=gen)
    a_hi, a_lo = operands(a.to(cuda), dtype)
    w_hi, w_lo = operands(wt.to(cuda), dtype)
    bias_dev = bias.to(cuda)
    odt = torch.bfloat16 if dtype == ffi.PB_BF16 else torch.float32
    out = torch.empty((m, d), dtype=odt, device=cuda)
    lib = ffi.lib()
    n_part = lib.pb_rgcn_gemm_fwd_bn_partial_rows(m)
    parts = torch.zeros((n_part, 2, d), device=cuda)
    ffi.check(lib.pb_rgcn_gemm_fwd_bn(ptr(a_hi), ptr(a_lo), k, ptr(w_hi), ptr(w_lo), ptr(bias_dev), ptr(out), d, m, d, k, gref,
                                      dtype, act, ptr(parts), st()), "gemm_fwd_bn")
    stored = out.float().double().cpu()[valid]
    torch.testing.assert_close(parts[:, 0].sum(0).double().cpu(), stored.sum(0), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(parts[:, 1].sum(0).double().cpu(), (stored ** 2).sum(0), rtol=1e-5, atol=1e-3)
    gamma, beta = (torch.rand(d, generator=gen) + 0.5).to(cuda), torch.randn(d, generator=gen).to(cuda)
    res = []
    for fused in (True, False):
        rm, rv = torch.zeros(d, device=cuda), torch.ones(d, device=cuda)
        save, coef = torch.empty(2, d, device=cuda), torch.empty(3, d, device=cuda)
        if fused:
            ffi.check(lib.pb_bn_finalize(ptr(parts), n_part, int(valid.sum()), d, ptr(gamma), ptr(beta), 1e-5, 0.1, ptr(rm), ptr(rv),
                                         ptr(save), ptr(coef), st()), "bn_finalize")
        else:
            ws_bytes = lib.pb_bn_workspace_bytes(m, d)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
            ffi.check(lib.pb_bn_stats(ptr(out), d, m, d, gref, ptr(gamma), ptr(beta), 1e-5, 0.1, ptr(rm), ptr(rv), ptr(save), ptr(coef),
                                      ptr(ws), ws_bytes, act, st()), "bn_stats")
        res.append((rm, rv, save, coef))
    for x, y in zip(*res):
        torch.testing.assert_close(x, y, rtol=2e-5, atol=2e-6)
    want_mean, want_var = stored.mean(0), stored.var(0, unbiased=False)
    torch.testing.assert_close(res[0][2][0].double().cpu(), want_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res[0][2][1].double().cpu(), 1 / torch.sqrt(want_var + 1e-5), rtol=1e-5, atol=1e-6)
