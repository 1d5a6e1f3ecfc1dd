"""Autograd bindings of the message-passing kernels (thin host glue over the C ABI).

One ``torch.autograd.Function`` covers a whole GCN layer of the reference (model.py:198-206):

    out = GCL(x)            model.py:55-121   pb_edge_table_fwd, pb_agg_fwd, pb_weight_prep, pb_rgcn_gemm_fwd
    y   = x_res + relu(BN(out))               pb_bn_stats, pb_bn_relu_res_fwd

and its backward (what autograd derives for those lines in the reference):

    pb_bn_relu_res_bwd -> pb_agg_fwd (operand recompute) -> pb_rgcn_gemm_bwd_weight, pb_rgcn_gemm_bwd_data
    -> pb_agg_bwd -> pb_edge_table_bwd

PyTorch only owns the memory and the stream here; no arithmetic of the path runs in ATen.
"""
from __future__ import annotations

import itertools
import os
import threading
from dataclasses import dataclass
from typing import Optional

import torch

from . import _ffi
from .graph import CsrPlan

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

_state = threading.local()
_PRECISIONS = {"fp32": _ffi.PB_F32, "bf16": _ffi.PB_BF16}
_default_precision = "fp32"
_seed_counter = itertools.count()

launch_counter = _ffi.launch_counter


_structured = os.environ.get("PB200_STRUCTURED", "1") != "0"


def set_structured(enabled: bool) -> None:
    """Use the track-relation-sorted 4d-wide operand layout inside GCN stacks when the graph allows it."""
    global _structured
    _structured = bool(enabled)


def structured_enabled() -> bool:
    return _structured


# content decoder: un-embed drum rows and the others separately (each with its own pitch head) when only the loss is needed
_split_heads = os.environ.get("PB200_SPLIT_HEADS", "1") != "0"


def split_heads_enabled() -> bool:
    return _split_heads


_bf16_activations = os.environ.get("PB200_BF16_ACT", "1") != "0"
# "fused" (default): scatter-by-source and the edge-table gradient in one pass, shared-memory accumulators with
# thread-owned columns; "legacy": the round-1 pair of kernels with the E x d intermediate (kept for A/B measurements)
_agg_bwd_mode = os.environ.get("PB200_AGG_BWD", "fused")
# BatchNorm batch statistics as a by-product of the forward GEMM's epilogue (default) or by pb_bn_stats' own pass over `out`
_bn_stats_in_epilogue = os.environ.get("PB200_BN_EPILOGUE", "1") != "0"


def set_bf16_activations(enabled: bool) -> None:
    """bf16 mode only: keep the activations of a structured GCN stack in bf16 between its kernels (default) or in fp32."""
    global _bf16_activations
    _bf16_activations = bool(enabled)


def bf16_activations_enabled() -> bool:
    return _bf16_activations


def set_precision(name: str) -> None:
    """'fp32' (TF32x3 tensor-core mode, fp32-grade) or 'bf16' (bf16 operands, fp32 accumulate)."""
    global _default_precision
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    _default_precision = name


def get_precision() -> str:
    return _default_precision


def _call(name: str, *args, tag=None) -> None:
    _ffi.call(name, *args, tag=tag)


def next_seed() -> int:
    """Per-call dropout seed: deterministic under torch.manual_seed and call order, no device sync."""
    base = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    return (base * 0x9E3779B97F4A7C15 + next(_seed_counter) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


@dataclass
class LayerConfig:
    dtype: int                 # _ffi.PB_F32 / PB_BF16
    p_drop: float              # GCL message dropout (model.py:133)
    seed: int
    training: bool
    batch_norm: bool           # fuse BN + ReLU + residual (GCN layer) or return the raw GCL output
    save_operand: bool = False  # keep the aggregated GEMM operand for backward instead of recomputing it
    keep_bits: object = None    # (bits, ready event) drawn ahead of the call by prefetch_keep_bits, else None
    eps: float = BN_EPS
    momentum: float = BN_MOMENTUM


def _operand(n: int, k: int, dtype: int, dev, zero: bool = False):
    """GEMM operand buffers in the arithmetic mode's storage: bf16, or TF32 hi/lo fp32 pair."""
    make = torch.zeros if zero else torch.empty
    if dtype == _ffi.PB_BF16:
        return make((n, k), dtype=torch.bfloat16, device=dev), None
    return make((n, k), dtype=torch.float32, device=dev), make((n, k), dtype=torch.float32, device=dev)


def _edge_table(nn_w: torch.Tensor, nn_b: torch.Tensor, d: int, st: int) -> torch.Tensor:
    table = torch.empty((_ffi.N_DISTS, d), dtype=torch.float32, device=nn_w.device)
    _call("pb_edge_table_fwd", nn_w.data_ptr(), nn_b.data_ptr(), d, table.data_ptr(), st)
    return table


def _weights(weight, root, r, d, dtype, st):
    dev = weight.device
    k = (r + 1) * d
    tdt = torch.bfloat16 if dtype == _ffi.PB_BF16 else torch.float32
    w_hi = torch.empty((k, d), dtype=tdt, device=dev)
    wt_hi = torch.empty((d, k), dtype=tdt, device=dev)
    w_lo = wt_lo = None
    if dtype == _ffi.PB_F32:
        w_lo, wt_lo = torch.empty_like(w_hi), torch.empty_like(wt_hi)
    _call("pb_weight_prep", weight.data_ptr(), root.data_ptr(), r, d, dtype, w_hi.data_ptr(), _ffi.ptr(w_lo),
          wt_hi.data_ptr(), _ffi.ptr(wt_lo), st)
    return w_hi, w_lo, wt_hi, wt_lo


def _keep_bits(n_edges: int, d: int, p: float, seed: int, dev, st: int):
    """Packed dropout keep-bits of one layer call (None when dropout is off)."""
    if p <= 0.0 or n_edges == 0:
        return None
    nbytes = _ffi.lib().pb_dropout_bits_bytes(n_edges, d)
    bits = torch.empty(nbytes // 2, dtype=torch.int16, device=dev)
    _call("pb_dropout_bits", n_edges, d, float(p), int(seed), bits.data_ptr(), st)
    return bits


_bits_prefetch = os.environ.get("PB200_BITS_PREFETCH", "1") != "0"
_bits_streams = {}


def prefetch_keep_bits(plan: CsrPlan, d: int, p: float, n_calls: int):
    """Draws the keep-bits of the next ``n_calls`` layer calls of a stack (same seeds, same order as the calls would
    draw them) on a side stream: pb_dropout_bits is pure integer arithmetic on no input, so it fills issue slots the
    tensor-core- and HBM-bound kernels of the preceding layers leave idle instead of sitting between them.
    Returns [(seed, (bits, ready_event))] or None when there is nothing to draw."""
    if not _bits_prefetch or p <= 0.0 or plan.n_edges == 0 or n_calls <= 0:
        return None
    dev = plan.device
    side = _bits_streams.get(dev)
    if side is None:
        side = _bits_streams[dev] = torch.cuda.Stream(device=dev)
    nbytes = _ffi.lib().pb_dropout_bits_bytes(plan.n_edges, d)
    main = torch.cuda.current_stream(dev)
    seeds = [next_seed() for _ in range(n_calls)]
    bufs = [torch.empty(nbytes // 2, dtype=torch.int16, device=dev) for _ in range(n_calls)]   # owned by `main`
    side.wait_stream(main)               # the buffers' previous users on `main` are done before the side stream writes
    out = []
    with _ffi.on_device(dev):
        for seed, bits in zip(seeds, bufs):
            _call("pb_dropout_bits", plan.n_edges, d, float(p), int(seed), bits.data_ptr(), side.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(side)
            out.append((seed, (bits, ev)))
    return out


class RGCLayerFn(torch.autograd.Function):
    """y = x_res + relu(BN(GCL(x)))   (cfg.batch_norm)   or   y = GCL(x)   (plain layer)."""

    @staticmethod
    def forward(ctx, x, weight, root, bias, nn_w, nn_b, gamma, beta, running_mean, running_var,
                plan: CsrPlan, cfg: LayerConfig, struct=None):
        _ffi.require_cuda(x, weight, root, nn_w, nn_b)
        # activation storage: fp32 (the API dtype) or, inside a bf16-mode GCN stack, bf16 (see GCN.forward)
        if x.dtype == torch.bfloat16:
            if cfg.dtype != _ffi.PB_BF16 or not cfg.batch_norm:
                raise TypeError("bf16 node features are only taken by BatchNorm-fused layers in the bf16 mode")
        elif x.dtype != torch.float32:
            raise TypeError("node features must be float32 (or bfloat16 inside a bf16-mode stack)")
        act = _ffi.PB_BF16 if x.dtype == torch.bfloat16 else _ffi.PB_F32
        x = x.contiguous()
        n, d = x.shape
        r = plan.n_relations                     # operand blocks per node: R, or 3 slots in the structured layout
        n_w = weight.shape[0]
        if weight.shape != (n_w, d, d) or root.shape != (d, d) or (struct is None and n_w != r):
            raise ValueError("the CUDA path supports in_channels == out_channels == d, weight [R,d,d], root [d,d]")
        if n != plan.n_nodes:
            raise ValueError(f"x has {n} rows but the graph has {plan.n_nodes} nodes")
        k = (r + 1) * d
        groups = None if struct is None else struct.groups_ref()
        dev = x.device
        weight, root = weight.contiguous(), root.contiguous()
        nn_w, nn_b = nn_w.contiguous(), nn_b.contiguous()
        bias_c = None if bias is None else bias.contiguous()
        with _ffi.on_device(dev):
            st = _ffi.stream()
            table = _edge_table(nn_w, nn_b, d, st)
            a_hi, a_lo = _operand(n, k, cfg.dtype, dev)
            p = cfg.p_drop if cfg.training else 0.0
            if cfg.keep_bits is not None and p > 0.0:
                keep_bits, ready = cfg.keep_bits
                torch.cuda.current_stream(dev).wait_event(ready)
            else:
                keep_bits = _keep_bits(plan.n_edges, d, p, cfg.seed, dev, st)
            cfg.keep_bits = None
            _call("pb_agg_fwd", plan.ref(), x.data_ptr(), d, table.data_ptr(), a_hi.data_ptr(), _ffi.ptr(a_lo), k,
                  cfg.dtype, _ffi.ptr(keep_bits), p, act, st)
            ctx.keep_bits = keep_bits
            # one operand preparation per layer call: the transposed copy feeds this GEMM, the other one the input
            # gradient in backward (kept in ctx: 3.7 MB at d = 512 in bf16)
            w_hi, w_lo, wt_hi, wt_lo = _weights(weight, root, n_w, d, cfg.dtype, st)
            ctx.w_operand = (w_hi, w_lo)
            ctx.table = table
            out = torch.empty((n, d), dtype=x.dtype, device=dev)
            fuse_stats = cfg.batch_norm and cfg.training and _bn_stats_in_epilogue
            if fuse_stats:       # BatchNorm column sums leave the GEMM epilogue: no second pass over `out`
                n_part = int(_ffi.lib().pb_rgcn_gemm_fwd_bn_partial_rows(n))
                bn_part = torch.zeros((n_part, 2, d), dtype=torch.float32, device=dev)
                _call("pb_rgcn_gemm_fwd_bn", a_hi.data_ptr(), _ffi.ptr(a_lo), k, wt_hi.data_ptr(), _ffi.ptr(wt_lo),
                      _ffi.ptr(bias_c), out.data_ptr(), d, n, d, k, groups, cfg.dtype, act, bn_part.data_ptr(), st)
            else:
                _call("pb_rgcn_gemm_fwd", a_hi.data_ptr(), _ffi.ptr(a_lo), k, wt_hi.data_ptr(), _ffi.ptr(wt_lo),
                      _ffi.ptr(bias_c), out.data_ptr(), d, n, d, k, groups, cfg.dtype, act, st)
            ctx.operand = (a_hi, a_lo) if cfg.save_operand else None
            ctx.struct = struct
            if not cfg.batch_norm:
                ctx.save_for_backward(x, weight, root, nn_w, nn_b)
                ctx.plan, ctx.cfg, ctx.has_bias = plan, cfg, bias is not None
                return out
            coef = torch.empty((3, d), dtype=torch.float32, device=dev)
            save = torch.empty((2, d), dtype=torch.float32, device=dev)
            if fuse_stats:
                n_valid = n if struct is None else sum(struct.counts)
                _call("pb_bn_finalize", bn_part.data_ptr(), n_part, n_valid, d, gamma.data_ptr(), beta.data_ptr(), cfg.eps,
                      cfg.momentum, _ffi.ptr(running_mean), _ffi.ptr(running_var), save.data_ptr(), coef.data_ptr(), st)
            elif cfg.training:
                ws_bytes = _ffi.lib().pb_bn_workspace_bytes(n, d)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _call("pb_bn_stats", out.data_ptr(), d, n, d, groups, gamma.data_ptr(), beta.data_ptr(), cfg.eps,
                      cfg.momentum, _ffi.ptr(running_mean), _ffi.ptr(running_var), save.data_ptr(), coef.data_ptr(),
                      ws.data_ptr(), ws_bytes, act, st)
            else:
                _call("pb_bn_prepare_eval", gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                      running_var.data_ptr(), cfg.eps, d, coef.data_ptr(), st)
            y = torch.empty((n, d), dtype=x.dtype, device=dev)
            _call("pb_bn_relu_res_fwd", out.data_ptr(), d, x.data_ptr(), coef.data_ptr(), y.data_ptr(), n, d, groups, 1,
                  act, st)
        ctx.save_for_backward(x, weight, root, nn_w, nn_b, gamma, out, coef, save)
        ctx.plan, ctx.cfg, ctx.has_bias = plan, cfg, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        cfg: LayerConfig = ctx.cfg
        plan: CsrPlan = ctx.plan
        if cfg.batch_norm and not cfg.training:
            raise NotImplementedError("backward through eval-mode BatchNorm is not part of the training path")
        if cfg.batch_norm:
            x, weight, root, nn_w, nn_b, gamma, out, coef, save = ctx.saved_tensors
        else:
            x, weight, root, nn_w, nn_b = ctx.saved_tensors
        act = _ffi.PB_BF16 if x.dtype == torch.bfloat16 else _ffi.PB_F32
        gy = gy.to(x.dtype).contiguous()             # the gradient travels in the activations' storage dtype
        n, d = x.shape
        r = plan.n_relations
        n_w = weight.shape[0]
        k = (r + 1) * d
        dev = x.device
        lib = _ffi.lib()
        struct = ctx.struct
        groups = None if struct is None else struct.groups_ref()
        with _ffi.on_device(dev):
            st = _ffi.stream()
            # padding rows of the structured layout must be zero: pb_bn_relu_res_bwd writes them itself
            g_hi, g_lo = _operand(n, d, cfg.dtype, dev, zero=struct is not None and not cfg.batch_norm)
            g_bias = torch.empty(d, dtype=torch.float32, device=dev)
            ws_bytes = lib.pb_bn_workspace_bytes(n, d)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            g_gamma = g_beta = None
            if cfg.batch_norm:
                g_gamma = torch.empty(d, dtype=torch.float32, device=dev)
                g_beta = torch.empty(d, dtype=torch.float32, device=dev)
                _call("pb_bn_relu_res_bwd", gy.data_ptr(), out.data_ptr(), d, gamma.data_ptr(), save.data_ptr(),
                      coef.data_ptr(), n, d, groups, cfg.dtype, g_hi.data_ptr(), _ffi.ptr(g_lo), d, g_gamma.data_ptr(),
                      g_beta.data_ptr(), g_bias.data_ptr(), ws.data_ptr(), ws_bytes, act, st)
            else:
                _call("pb_grad_prep", gy.data_ptr(), d, n, d, cfg.dtype, g_hi.data_ptr(), _ffi.ptr(g_lo), d,
                      g_bias.data_ptr(), ws.data_ptr(), ws_bytes, st)
            table = ctx.table
            p = cfg.p_drop if cfg.training else 0.0
            if ctx.operand is not None:
                a_hi, a_lo = ctx.operand
                ctx.operand = None
            else:  # recompute the aggregated operand instead of keeping N x (R+1)d per layer alive
                a_hi, a_lo = _operand(n, k, cfg.dtype, dev)
                _call("pb_agg_fwd", plan.ref(), x.data_ptr(), d, table.data_ptr(), a_hi.data_ptr(), _ffi.ptr(a_lo), k,
                      cfg.dtype, _ffi.ptr(ctx.keep_bits), p, act, st)
            if ctx.w_operand is not None:
                w_hi, w_lo = ctx.w_operand
                ctx.w_operand = None
            else:
                w_hi, w_lo, _, _ = _weights(weight, root, n_w, d, cfg.dtype, st)
            kw = (n_w + 1) * d
            # structured layout: one split-K launch whose splits follow the row groups + a grouped reduce that sums
            # the track block per group (weight[0..3]) and the onset / next / root blocks over all rows
            d_wcat = torch.empty((kw, d), dtype=torch.float32, device=dev)
            wws_bytes = lib.pb_rgcn_gemm_bwd_weight_workspace_bytes_for(n, d, k, cfg.dtype)
            wws = torch.empty(wws_bytes, dtype=torch.uint8, device=dev)
            _call("pb_rgcn_gemm_bwd_weight", a_hi.data_ptr(), _ffi.ptr(a_lo), k, g_hi.data_ptr(), _ffi.ptr(g_lo), d,
                  d_wcat.data_ptr(), n, d, k, groups, cfg.dtype, wws.data_ptr(), wws_bytes, st)
            del a_hi, a_lo, wws
            d_a = torch.empty((n, k), dtype=torch.bfloat16 if cfg.dtype == _ffi.PB_BF16 else torch.float32, device=dev)
            _call("pb_rgcn_gemm_bwd_data", g_hi.data_ptr(), _ffi.ptr(g_lo), d, w_hi.data_ptr(), _ffi.ptr(w_lo),
                  d_a.data_ptr(), k, n, d, k, groups, cfg.dtype, st)
            gx = torch.empty((n, d), dtype=x.dtype, device=dev)
            g_nn_w = torch.empty((d, _ffi.N_DISTS), dtype=torch.float32, device=dev)
            g_nn_b = torch.empty(d, dtype=torch.float32, device=dev)
            gy_res = gy.data_ptr() if cfg.batch_norm else None
            if _agg_bwd_mode == "fused":
                n_part = int(lib.pb_agg_bwd_num_partials(n, d, cfg.dtype))
                partials = torch.empty((n_part, _ffi.N_DISTS, d), dtype=torch.float32, device=dev)
                _call("pb_agg_bwd_fused", plan.ref(), x.data_ptr(), d, table.data_ptr(), d_a.data_ptr(), k, cfg.dtype,
                      gy_res, gx.data_ptr(), partials.data_ptr(), _ffi.ptr(ctx.keep_bits), p, act, st)
                _call("pb_edge_table_bwd_fused", partials.data_ptr(), n_part, d, g_nn_w.data_ptr(), g_nn_b.data_ptr(), st)
            else:
                q_buf = torch.empty((max(plan.n_edges, 1), d), dtype=d_a.dtype, device=dev)
                partials = torch.empty((plan.n_dist_items, d), dtype=torch.float32, device=dev)
                _call("pb_agg_bwd", plan.ref(), x.data_ptr(), d, table.data_ptr(), d_a.data_ptr(), k, cfg.dtype,
                      gy_res, gx.data_ptr(), q_buf.data_ptr(), partials.data_ptr(), _ffi.ptr(ctx.keep_bits), p, act, st)
                _call("pb_edge_table_bwd", partials.data_ptr(), plan.dist_item_ptr.data_ptr(), d, g_nn_w.data_ptr(),
                      g_nn_b.data_ptr(), st)
        g_weight = d_wcat[: n_w * d].view(n_w, d, d)
        g_root = d_wcat[n_w * d:]
        return (gx, g_weight, g_root, g_bias if ctx.has_bias else None, g_nn_w, g_nn_b, g_gamma, g_beta,
                None, None, None, None, None)


# Operands up to this size are kept for backward (LMD16, per-GPU batch 256, bf16: 0.94 GB per layer); larger ones
# (big batches, the fp32 hi/lo mode) are recomputed from x so that activation memory stays ~2 N d floats per layer.
SAVE_OPERAND_MAX_BYTES = int(os.environ.get("PB200_SAVE_OPERAND_MB", "2048")) << 20


def rgc_layer(x, weight, root, bias, nn_w, nn_b, plan: CsrPlan, *, gamma=None, beta=None, running_mean=None,
              running_var=None, batch_norm: bool, training: bool, p_drop: float, precision: Optional[str] = None,
              seed: Optional[int] = None, eps: float = BN_EPS, momentum: float = BN_MOMENTUM,
              save_operand: Optional[bool] = None, struct=None, keep_bits=None):
    dtype = _PRECISIONS[precision or _default_precision]
    if save_operand is None:
        elem = 2 if dtype == _ffi.PB_BF16 else 8
        save_operand = (torch.is_grad_enabled() and
                        x.size(0) * (plan.n_relations + 1) * x.size(1) * elem <= SAVE_OPERAND_MAX_BYTES)
    cfg = LayerConfig(dtype=dtype, p_drop=float(p_drop),
                      seed=next_seed() if seed is None else int(seed), training=bool(training),
                      batch_norm=bool(batch_norm), save_operand=bool(save_operand), eps=float(eps),
                      momentum=float(momentum), keep_bits=keep_bits)
    return RGCLayerFn.apply(x, weight, root, bias, nn_w, nn_b, gamma, beta, running_mean, running_var, plan, cfg,
                            struct)


# ----------------------------------------------------------------------------------------------------------------
# Dense layers adjacent to the path (chord encoder / decoder) on the same tcgen05 GEMM.
def _tf32_split(t: torch.Tensor):
    t = t.contiguous()                                                        # lo must share hi's row-major layout
    hi = (t.view(torch.int32) & -8192).view(torch.float32)                    # top 19 bits: a valid TF32 value
    return hi, t - hi


def _as_operand(t: torch.Tensor, dtype: int):
    if dtype == _ffi.PB_BF16:
        return (t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)).contiguous(), None
    return _tf32_split(t.float())


class TensorCoreLinearFn(torch.autograd.Function):
    """y = x @ weight.T + bias with pb_gemm_nt (forward, input gradient) and the deterministic split-K
    pb_rgcn_gemm_bwd_weight (weight gradient). Precision follows the path's mode (bf16, or TF32x3 = fp32-grade)."""

    @staticmethod
    def forward(ctx, x, weight, bias, dtype: int, out_bf16: bool):
        m, k = x.shape
        n = weight.shape[0]
        dev = x.device
        x_hi, x_lo = _as_operand(x, dtype)
        w_hi, w_lo = _as_operand(weight, dtype)
        out_bf16 = bool(out_bf16 and dtype == _ffi.PB_BF16)
        out = torch.empty((m, n), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
        bias_f = None if bias is None else bias.float().contiguous()
        with _ffi.on_device(dev):
            _call("pb_gemm_nt", x_hi.data_ptr(), _ffi.ptr(x_lo), k, w_hi.data_ptr(), _ffi.ptr(w_lo), k,
                  _ffi.ptr(bias_f), out.data_ptr(), n, m, n, k, dtype, int(out_bf16), _ffi.stream(), tag="linear")
        ctx.save_for_backward(x_hi, x_lo, weight)
        ctx.dtype, ctx.has_bias, ctx.x_dtype = dtype, bias is not None, x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        x_hi, x_lo, weight = ctx.saved_tensors
        dtype = ctx.dtype
        m, k = x_hi.shape
        n = weight.shape[0]
        dev = g.device
        g_hi, g_lo = _as_operand(g, dtype)
        lib = _ffi.lib()
        with _ffi.on_device(dev):
            st = _ffi.stream()
            dx = None
            if ctx.needs_input_grad[0]:
                wt_hi, wt_lo = _as_operand(weight.t(), dtype)                   # [k, n], contraction dim contiguous
                dx_bf16 = dtype == _ffi.PB_BF16 and ctx.x_dtype == torch.bfloat16
                dx = torch.empty((m, k), dtype=torch.bfloat16 if dx_bf16 else torch.float32, device=dev)
                _call("pb_gemm_nt", g_hi.data_ptr(), _ffi.ptr(g_lo), n, wt_hi.data_ptr(), _ffi.ptr(wt_lo), n, None,
                      dx.data_ptr(), k, m, k, n, dtype, int(dx_bf16), st, tag="linear")
            dw = None
            if ctx.needs_input_grad[1]:
                dw = torch.empty((n, k), dtype=torch.float32, device=dev)
                ws_bytes = lib.pb_rgcn_gemm_bwd_weight_workspace_bytes_for(m, k, n, dtype)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                # dW[n, k] = g^T @ x: the split-K kernel contracts over the rows of its two [m, .] operands
                _call("pb_rgcn_gemm_bwd_weight", g_hi.data_ptr(), _ffi.ptr(g_lo), n, x_hi.data_ptr(), _ffi.ptr(x_lo), k,
                      dw.data_ptr(), m, k, n, None, dtype, ws.data_ptr(), ws_bytes, st, tag="linear")
        db = g.sum(0, dtype=torch.float32) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


class SplitRowsLinearFn(torch.autograd.Function):
    """Two Linears over the two row blocks of one matrix: rows [0, n0) of x go through (w0, b0), the rest through
    (w1, b1); both share the contraction width. Used by the content decoder's un-embedding (drum rows first, then the
    others: each row only gets the pitch head of its own instrument class). One input-gradient buffer is written by
    both blocks (no per-slice zero-fill + add in autograd), weight gradients by the split-K kernel."""

    @staticmethod
    def forward(ctx, x, n0: int, w0, b0, w1, b1, dtype: int, out_bf16: bool):
        m, k = x.shape
        dev = x.device
        x_hi, x_lo = _as_operand(x, dtype)
        out_bf16 = bool(out_bf16 and dtype == _ffi.PB_BF16)
        odt = torch.bfloat16 if out_bf16 else torch.float32
        esz = x_hi.element_size()
        outs, saved_w = [], []
        with _ffi.on_device(dev):
            for r0, rows, w, b in ((0, n0, w0, b0), (n0, m - n0, w1, b1)):
                n = w.shape[0]
                out = torch.empty((rows, n), dtype=odt, device=dev)
                if rows > 0:
                    w_hi, w_lo = _as_operand(w, dtype)
                    bias_f = b.float().contiguous()
                    _call("pb_gemm_nt", x_hi.data_ptr() + r0 * k * esz, None if x_lo is None else x_lo.data_ptr() + r0 * k * esz,
                          k, w_hi.data_ptr(), _ffi.ptr(w_lo), k, bias_f.data_ptr(), out.data_ptr(), n, rows, n, k, dtype,
                          int(out_bf16), _ffi.stream(), tag="linear")
                outs.append(out)
        ctx.save_for_backward(x_hi, x_lo, w0, w1)
        ctx.dtype, ctx.n0, ctx.x_dtype = dtype, n0, x.dtype
        return tuple(outs)

    @staticmethod
    def backward(ctx, g0, g1):
        x_hi, x_lo, w0, w1 = ctx.saved_tensors
        dtype, n0 = ctx.dtype, ctx.n0
        m, k = x_hi.shape
        dev = x_hi.device
        lib = _ffi.lib()
        esz = x_hi.element_size()
        dx_bf16 = dtype == _ffi.PB_BF16 and ctx.x_dtype == torch.bfloat16
        dx = torch.empty((m, k), dtype=torch.bfloat16 if dx_bf16 else torch.float32, device=dev)
        grads = []
        with _ffi.on_device(dev):
            st = _ffi.stream()
            for r0, rows, w, g in ((0, n0, w0, g0), (n0, m - n0, w1, g1)):
                n = w.shape[0]
                if rows == 0 or g is None:
                    dx[r0:r0 + rows].zero_()
                    grads += [torch.zeros_like(w), torch.zeros(n, dtype=torch.float32, device=dev)]
                    continue
                g_hi, g_lo = _as_operand(g, dtype)
                wt_hi, wt_lo = _as_operand(w.t(), dtype)
                _call("pb_gemm_nt", g_hi.data_ptr(), _ffi.ptr(g_lo), n, wt_hi.data_ptr(), _ffi.ptr(wt_lo), n, None,
                      dx.data_ptr() + r0 * k * dx.element_size(), k, rows, k, n, dtype, int(dx_bf16), st, tag="linear")
                dw = torch.empty((n, k), dtype=torch.float32, device=dev)
                ws_bytes = lib.pb_rgcn_gemm_bwd_weight_workspace_bytes_for(rows, k, n, dtype)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _call("pb_rgcn_gemm_bwd_weight", g_hi.data_ptr(), _ffi.ptr(g_lo), n, x_hi.data_ptr() + r0 * k * esz,
                      None if x_lo is None else x_lo.data_ptr() + r0 * k * esz, k, dw.data_ptr(), rows, k, n, None, dtype,
                      ws.data_ptr(), ws_bytes, st, tag="linear")
                grads += [dw, g.sum(0, dtype=torch.float32)]
        return dx, None, grads[0], grads[1], grads[2], grads[3], None, None


def split_rows_linear(x, n0: int, w0, b0, w1, b1, out_bf16: bool = False, precision: Optional[str] = None):
    """(x[:n0] @ w0.T + b0, x[n0:] @ w1.T + b1) on the tcgen05 GEMM (x 2-D CUDA, widths multiples of 64)."""
    dtype = _PRECISIONS[precision or _default_precision]
    return SplitRowsLinearFn.apply(x, int(n0), w0, b0, w1, b1, dtype, out_bf16)


class TableGatherFn(torch.autograd.Function):
    """out = table[ids] for a tiny table and millions of ids (the token-table embeddings of vae.ContentEncoder).

    The gradient of such a gather is dTable = onehot(ids)^T @ g: a [vocab x rows] by [rows x c] contraction over the
    rows — exactly the deterministic split-K weight-gradient GEMM of the path, instead of the library's sort-based
    scatter (1.8 ms per table at 2M ids)."""

    @staticmethod
    def forward(ctx, table, ids, dtype: int):
        ctx.save_for_backward(ids)
        ctx.vocab, ctx.table_dtype, ctx.dtype = table.size(0), table.dtype, dtype
        return torch.nn.functional.embedding(ids, table)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        dtype, vocab = ctx.dtype, ctx.vocab
        c = g.size(-1)
        m = ids.numel()
        vp = (vocab + 63) // 64 * 64
        dev = g.device
        onehot = torch.zeros((m, vp), dtype=torch.bfloat16 if dtype == _ffi.PB_BF16 else torch.float32, device=dev)
        onehot.scatter_(1, ids.reshape(-1, 1), 1.0)
        g_hi, g_lo = _as_operand(g.reshape(m, c), dtype)
        oh_lo = None if dtype == _ffi.PB_BF16 else torch.zeros_like(onehot)     # 0/1 is exact in TF32
        lib = _ffi.lib()
        d_table = torch.empty((vp, c), dtype=torch.float32, device=dev)
        ws_bytes = lib.pb_rgcn_gemm_bwd_weight_workspace_bytes_for(m, c, vp, dtype)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _ffi.on_device(dev):
            _call("pb_rgcn_gemm_bwd_weight", onehot.data_ptr(), _ffi.ptr(oh_lo), vp, g_hi.data_ptr(), _ffi.ptr(g_lo), c,
                  d_table.data_ptr(), m, c, vp, None, dtype, ws.data_ptr(), ws_bytes, _ffi.stream(), tag="linear")
        return d_table[:vocab].to(ctx.table_dtype), None, None


def table_gather(table: torch.Tensor, ids: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
    """table[ids] with the tensor-core gradient above (CUDA, channel count a multiple of 64); F.embedding otherwise."""
    if not (table.is_cuda and table.size(1) % 64 == 0 and torch.is_grad_enabled() and table.requires_grad):
        return torch.nn.functional.embedding(ids, table)
    return TableGatherFn.apply(table, ids, _PRECISIONS[precision or _default_precision])


def tc_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], out_bf16: bool = False,
              precision: Optional[str] = None) -> torch.Tensor:
    """nn.Linear on the tcgen05 GEMM for 2-D CUDA inputs whose sizes are multiples of 64; plain F.linear
    otherwise (CPU construction / odd shapes are outside the accelerated path, not a fallback of it)."""
    n, k = weight.shape
    if not (x.is_cuda and x.dim() == 2 and n % 64 == 0 and k % 64 == 0):
        return torch.nn.functional.linear(x, weight, bias)
    dtype = _PRECISIONS[precision or _default_precision]
    return TensorCoreLinearFn.apply(x, weight, bias, dtype, out_bf16)


def _act_code(dtype: torch.dtype) -> int:
    return _ffi.PB_BF16 if dtype == torch.bfloat16 else _ffi.PB_F32


def _rows_scatter(x, pos, struct, out_dtype):
    out = torch.empty((struct.n_padded, x.size(1)), dtype=out_dtype, device=x.device)
    with _ffi.on_device(x.device):
        _call("pb_rows_scatter", x.data_ptr(), _act_code(x.dtype), pos.data_ptr(), x.size(0), x.size(1), out.data_ptr(),
              _act_code(out_dtype), struct.n_padded, struct.groups_ref(), _ffi.stream())
    return out


def _rows_gather(xp, pos, out_dtype):
    out = torch.empty((pos.numel(), xp.size(1)), dtype=out_dtype, device=xp.device)
    with _ffi.on_device(xp.device):
        _call("pb_rows_gather", xp.data_ptr(), _act_code(xp.dtype), pos.data_ptr(), pos.numel(), xp.size(1), out.data_ptr(),
              _act_code(out_dtype), _ffi.stream())
    return out


class ScatterRowsFn(torch.autograd.Function):
    """Entry of a structured GCN stack: out = zeros(n_padded, d) in ``out_dtype``; out[pos] = x (node -> padded row of
    the structured layout, injective). One kernel incl. the fp32 -> bf16 storage conversion and the zeroed padding rows
    (pb_rows_scatter); the gradient is the matching gather (autograd's own formula for index_copy goes through
    index_add_ with atomic adds, which an injective map does not need)."""

    @staticmethod
    def forward(ctx, x, struct, out_dtype):
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        ctx.struct, ctx.x_dtype = struct, x.dtype
        return _rows_scatter(x.contiguous(), struct.pos, struct, out_dtype)

    @staticmethod
    def backward(ctx, g):
        return _rows_gather(g.contiguous(), ctx.struct.pos, ctx.x_dtype), None, None


class GatherRowsFn(torch.autograd.Function):
    """Exit of a structured GCN stack: out = xp[pos] in ``out_dtype`` (pb_rows_gather); the gradient scatters back
    (rows nobody read — the padding — get zero)."""

    @staticmethod
    def forward(ctx, xp, struct, out_dtype):
        ctx.struct, ctx.xp_dtype = struct, xp.dtype
        return _rows_gather(xp.contiguous(), struct.pos, out_dtype)

    @staticmethod
    def backward(ctx, g):
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        return _rows_scatter(g.contiguous(), ctx.struct.pos, ctx.struct, ctx.xp_dtype), None, None


class BarPoolFn(torch.autograd.Function):
    """out[b] = sum over the bar's nodes of softmax(gate)_v * h[v] (pb_bar_pool_fwd / pb_bar_pool_bwd)."""

    @staticmethod
    def forward(ctx, h, gate, bar_ptr):
        n, d = h.shape
        n_bars = bar_ptr.numel() - 1
        h, gate = h.float().contiguous(), gate.float().contiguous().view(-1)
        alpha = torch.zeros(n, dtype=torch.float32, device=h.device)
        out = torch.empty((n_bars, d), dtype=torch.float32, device=h.device)
        with _ffi.on_device(h.device):
            _call("pb_bar_pool_fwd", h.data_ptr(), d, gate.data_ptr(), bar_ptr.data_ptr(), n_bars, d, alpha.data_ptr(),
                  out.data_ptr(), _ffi.stream())
        ctx.save_for_backward(h, alpha, bar_ptr)
        return out

    @staticmethod
    def backward(ctx, g_out):
        h, alpha, bar_ptr = ctx.saved_tensors
        n, d = h.shape
        n_bars = bar_ptr.numel() - 1
        g_out = g_out.float().contiguous()
        g_h = torch.empty_like(h)
        g_gate = torch.empty(n, dtype=torch.float32, device=h.device)
        with _ffi.on_device(h.device):
            _call("pb_bar_pool_bwd", h.data_ptr(), d, alpha.data_ptr(), bar_ptr.data_ptr(), n_bars, d, g_out.data_ptr(),
                  g_h.data_ptr(), d, g_gate.data_ptr(), _ffi.stream())
        return g_h, g_gate, None


class BarExpandFn(torch.autograd.Function):
    """x[v] = z[bar(v)] with the deterministic segment-sum gradient (pb_bar_expand_fwd / pb_bar_expand_bwd)."""

    @staticmethod
    def forward(ctx, z, bar_ptr, n_nodes: int):
        n_bars, d = z.shape
        z = z.float().contiguous()
        x = torch.empty((n_nodes, d), dtype=torch.float32, device=z.device)
        with _ffi.on_device(z.device):
            _call("pb_bar_expand_fwd", z.data_ptr(), bar_ptr.data_ptr(), n_bars, d, x.data_ptr(), d, _ffi.stream())
        ctx.save_for_backward(bar_ptr)
        ctx.shape = (n_bars, d)
        return x

    @staticmethod
    def backward(ctx, g_x):
        (bar_ptr,) = ctx.saved_tensors
        n_bars, d = ctx.shape
        g_x = g_x.float().contiguous()
        g_z = torch.empty((n_bars, d), dtype=torch.float32, device=g_x.device)
        with _ffi.on_device(g_x.device):
            _call("pb_bar_expand_bwd", g_x.data_ptr(), d, bar_ptr.data_ptr(), n_bars, d, g_z.data_ptr(), _ffi.stream())
        return g_z, None, None


def _check_bar_ptr(bar_ptr: torch.Tensor, n_bars: int) -> None:
    if not (bar_ptr.is_cuda and bar_ptr.dtype == torch.int32 and bar_ptr.is_contiguous() and bar_ptr.numel() == n_bars + 1):
        raise ValueError("bar_ptr must be a contiguous CUDA int32 tensor of n_bars + 1 node offsets")


MAX_BAR_NODES = 128      # 4 tracks x 32 timesteps; pb_bar_pool_* keep a bar's scores in registers


def bar_pool(h: torch.Tensor, gate: torch.Tensor, bar_ptr: torch.Tensor) -> torch.Tensor:
    """Attention pooling of node rows h [N, d] per bar with scores gate [N]: f32 [n_bars, d]. Segments longer than 128
    nodes are not bars of this model: rejected by a device-side assertion (no host sync)."""
    _check_bar_ptr(bar_ptr, bar_ptr.numel() - 1)
    if bar_ptr.numel() > 1:
        torch._assert_async(((bar_ptr[1:] - bar_ptr[:-1]) <= MAX_BAR_NODES).all())
    return BarPoolFn.apply(h, gate.reshape(-1), bar_ptr)


def bar_expand(z: torch.Tensor, bar_ptr: torch.Tensor, n_nodes: int) -> torch.Tensor:
    """Rows z [n_bars, d] repeated over the nodes of their bar: f32 [n_nodes, d]."""
    _check_bar_ptr(bar_ptr, z.size(0))
    return BarExpandFn.apply(z, bar_ptr, n_nodes)


class ChordEmbedFn(torch.autograd.Function):
    """chord[v] = relu(bias + sum of the node's 30 folded-table rows): pb_chord_embed_fwd; the table gradient is
    onehot^T @ (g * relu') on the split-K weight-gradient GEMM (pb_chord_embed_bwd_prep builds both operands)."""

    @staticmethod
    def forward(ctx, tables, bias, tokens, set_id, tok_offset: int, dur_off: int, dtype: int):
        _, n_slots, vocab, d = tables.shape
        n = tokens.size(0)
        tab = (tables.to(torch.bfloat16) if dtype == _ffi.PB_BF16 else tables.float()).contiguous()
        bias_f = bias.float().contiguous()
        out = torch.empty((n, d), dtype=torch.float32, device=tables.device)
        with _ffi.on_device(tables.device):
            _call("pb_chord_embed_fwd", tokens.data_ptr(), tokens.stride(0), tok_offset, n_slots, set_id.data_ptr(),
                  tab.data_ptr(), dtype, vocab, dur_off, d, bias_f.data_ptr(), out.data_ptr(), d, n, _ffi.stream())
        ctx.save_for_backward(tokens, set_id, out)
        ctx.cfg = (n_slots, vocab, d, tok_offset, dur_off, dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        tokens, set_id, out = ctx.saved_tensors
        n_slots, vocab, d, tok_offset, dur_off, dtype = ctx.cfg
        n, dev = tokens.size(0), g.device
        bf16 = dtype == _ffi.PB_BF16
        vp = (vocab + 63) // 64 * 64
        kk = n_slots * vp
        g = g.float().contiguous()
        op_dtype = torch.bfloat16 if bf16 else torch.float32
        onehot = torch.empty((n, kk), dtype=op_dtype, device=dev)
        onehot_lo = None if bf16 else torch.zeros((n, kk), dtype=op_dtype, device=dev)      # 0/1 is exact in TF32
        gcat_hi = torch.empty((n, 2 * d), dtype=op_dtype, device=dev)
        gcat_lo = None if bf16 else torch.empty((n, 2 * d), dtype=op_dtype, device=dev)
        d_t = torch.empty((kk, 2 * d), dtype=torch.float32, device=dev)
        lib = _ffi.lib()
        ws_bytes = lib.pb_rgcn_gemm_bwd_weight_workspace_bytes_for(n, 2 * d, kk, dtype)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _ffi.on_device(dev):
            st = _ffi.stream()
            _call("pb_chord_embed_bwd_prep", tokens.data_ptr(), tokens.stride(0), tok_offset, n_slots, set_id.data_ptr(),
                  dur_off, vp, d, out.data_ptr(), d, g.data_ptr(), d, dtype, onehot.data_ptr(), gcat_hi.data_ptr(),
                  _ffi.ptr(gcat_lo), n, st)
            _call("pb_rgcn_gemm_bwd_weight", onehot.data_ptr(), _ffi.ptr(onehot_lo), kk, gcat_hi.data_ptr(),
                  _ffi.ptr(gcat_lo), 2 * d, d_t.data_ptr(), n, 2 * d, kk, None, dtype, ws.data_ptr(), ws_bytes, st,
                  tag="linear")
        d_tables = d_t.view(n_slots, vp, 2, d)[:, :vocab].permute(2, 0, 1, 3)
        if bf16:     # the prepared operand already holds the ReLU-masked gradient, one half per table set
            d_bias = gcat_hi.view(n, 2, d).sum(0, dtype=torch.float32).sum(0)
        else:
            d_bias = torch.where(out > 0, g, torch.zeros((), dtype=g.dtype, device=dev)).sum(0)
        return d_tables, d_bias, None, None, None, None, None


def chord_embed(tables: torch.Tensor, bias: torch.Tensor, tokens: torch.Tensor, set_id: torch.Tensor,
                tok_offset: int, dur_off: int, precision: Optional[str] = None) -> torch.Tensor:
    """Folded chord embedding (see csrc/chord.cu): tables f32 [2, n_slots, vocab, d], tokens int16 [N, stride] with
    (pitch, duration) pairs from ``tok_offset``, set_id uint8/bool [N] (1 = drum tables) -> relu'd chords f32 [N, d]."""
    if not (tables.is_cuda and tokens.dtype == torch.int16 and tokens.stride(-1) == 1 and set_id.is_contiguous()):
        raise ValueError("chord_embed needs CUDA tables, int16 tokens with contiguous rows and a contiguous set flag")
    tok2d = tokens.view(tokens.size(0), -1)
    flags = set_id.view(torch.uint8) if set_id.dtype == torch.bool else set_id
    return ChordEmbedFn.apply(tables, bias, tok2d, flags, tok_offset, dur_off, _PRECISIONS[precision or _default_precision])


class TokenNllFn(torch.autograd.Function):
    """Per-row negative log-likelihood of ``target`` under softmax(logits): pb_ce_fwd / pb_ce_bwd. Rows whose
    target equals ``ignore_index`` give 0 and no gradient (nn.CrossEntropyLoss(ignore_index=...), training.py:100)."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index: int):
        rows, classes = logits.shape
        dtype = _ffi.PB_BF16 if logits.dtype == torch.bfloat16 else _ffi.PB_F32
        nll = torch.empty(rows, dtype=torch.float32, device=logits.device)
        lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
        with _ffi.on_device(logits.device):
            _call("pb_ce_fwd", logits.data_ptr(), logits.stride(0), dtype, rows, classes, target.data_ptr(),
                  int(ignore_index), nll.data_ptr(), lse.data_ptr(), _ffi.stream())
        ctx.save_for_backward(logits, target, lse)
        ctx.ignore_index, ctx.dtype = int(ignore_index), dtype
        return nll

    @staticmethod
    def backward(ctx, g):
        logits, target, lse = ctx.saved_tensors
        rows, classes = logits.shape
        g = g.float().contiguous()
        grad = torch.empty((rows, classes), dtype=logits.dtype, device=logits.device)
        with _ffi.on_device(logits.device):
            _call("pb_ce_bwd", logits.data_ptr(), logits.stride(0), ctx.dtype, rows, classes, target.data_ptr(),
                  ctx.ignore_index, lse.data_ptr(), g.data_ptr(), grad.data_ptr(), classes, _ffi.stream())
        return grad, None, None


def _segment_args(specs, targets):
    """Host-side argument arrays of pb_ce_rows_*: widths, device target pointers, ignored ids."""
    import ctypes

    n = len(specs)
    widths = (ctypes.c_int32 * n)(*[w for _, w, _ in specs])
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in targets])
    ignore = (ctypes.c_int32 * n)(*[i for _, _, i in specs])
    return widths, ptrs, ignore


class TokenNllSegmentsFn(torch.autograd.Function):
    """TokenNllFn for several heads that share one logits matrix: head i owns the column block
    [col0_i, col0_i + width_i) of ``logits`` [rows, cols] and has its own targets / ignored id. One pass over each row
    per direction (pb_ce_rows_fwd / pb_ce_rows_bwd); the blocks tile the columns, so the backward fills ONE gradient
    matrix (no per-head slices to zero-fill and add up)."""

    @staticmethod
    def forward(ctx, logits, specs, *targets):
        rows, n = logits.size(0), len(specs)
        dtype = _ffi.PB_BF16 if logits.dtype == torch.bfloat16 else _ffi.PB_F32
        nll = torch.empty((n, rows), dtype=torch.float32, device=logits.device)
        lse = torch.empty((n, rows), dtype=torch.float32, device=logits.device)
        widths, ptrs, ignore = _segment_args(specs, targets)
        with _ffi.on_device(logits.device):
            _call("pb_ce_rows_fwd", logits.data_ptr(), logits.stride(0), dtype, rows, n, widths, ptrs, ignore,
                  nll.data_ptr(), lse.data_ptr(), _ffi.stream())
        ctx.save_for_backward(logits, lse, *targets)
        ctx.specs, ctx.dtype = specs, dtype
        return tuple(nll.unbind(0))

    @staticmethod
    def backward(ctx, *gs):
        n = len(ctx.specs)
        logits, lse, targets = ctx.saved_tensors[0], ctx.saved_tensors[1], ctx.saved_tensors[2:]
        rows, cols = logits.shape
        row_grad = torch.stack([torch.zeros(rows, dtype=torch.float32, device=logits.device) if g is None else g.float()
                                for g in gs]).contiguous()
        grad = torch.empty((rows, cols), dtype=logits.dtype, device=logits.device)
        widths, ptrs, ignore = _segment_args(ctx.specs, targets)
        with _ffi.on_device(logits.device):
            _call("pb_ce_rows_bwd", logits.data_ptr(), logits.stride(0), ctx.dtype, rows, n, widths, ptrs, ignore,
                  lse.data_ptr(), row_grad.data_ptr(), grad.data_ptr(), cols, _ffi.stream())
        return (grad, None) + (None,) * n


def token_nll_segments(logits: torch.Tensor, segments) -> tuple:
    """segments = [(col0, width, target int32 [rows], ignore_index), ...] tiling the columns of the CUDA ``logits``
    [rows, cols] (bf16 / fp32, unit column stride) -> one nll f32 [rows] per segment."""
    if not (logits.is_cuda and logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in (torch.bfloat16, torch.float32)):
        raise ValueError("token_nll_segments needs 2-D CUDA bf16/fp32 logits with contiguous rows")
    cover = 0
    for col0, width, target, _ in segments:
        if col0 != cover:
            raise ValueError("segments must tile the columns in order")
        if target.dtype != torch.int32 or not target.is_contiguous() or target.numel() != logits.size(0):
            raise ValueError("token_nll_segments needs one contiguous int32 target per row and segment")
        cover += width
    if cover != logits.size(1):
        raise ValueError(f"segments cover {cover} of {logits.size(1)} columns")
    specs = tuple((int(c), int(w), int(i)) for c, w, _, i in segments)
    return TokenNllSegmentsFn.apply(logits, specs, *[t for _, _, t, _ in segments])


def token_nll(logits: torch.Tensor, target: torch.Tensor, ignore_index: int) -> torch.Tensor:
    """nll f32 [rows] for CUDA logits [rows, classes] (bf16 or fp32, unit column stride) and int32 targets."""
    if not (logits.is_cuda and logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in (torch.bfloat16, torch.float32)):
        raise ValueError("token_nll needs 2-D CUDA bf16/fp32 logits with contiguous rows")
    if target.dtype != torch.int32 or not target.is_contiguous() or target.numel() != logits.size(0):
        raise ValueError("token_nll needs one contiguous int32 target per row")
    return TokenNllFn.apply(logits, target, ignore_index)


def dropout_keep_mask(n_edges: int, d: int, p_drop: float, seed: int, device) -> torch.Tensor:
    """The keep-mask pb_agg_fwd/bwd use for (seed, p): bool [E, d], indexed by edge_index column."""
    keep = torch.empty((n_edges, d), dtype=torch.uint8, device=device)
    with _ffi.on_device(device):
        _call("pb_dropout_mask", n_edges, d, float(p_drop), int(seed), keep.data_ptr(), _ffi.stream())
    return keep.bool()
