"""ctypes binding of libpolyphemus_b200.so (C ABI declared in include/polyphemus_b200.h).

There is no CPU fallback: if the shared library is missing, or a call fails, an exception is raised.
Build with ``python -m polyphemus_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# PB200_LIB: load another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(_PKG_DIR, "lib", "libpolyphemus_b200.so")

PB_F32 = 0
PB_BF16 = 1
N_RELATIONS = 6
N_DISTS = 32


class PolyphemusB200Error(RuntimeError):
    pass


class CsrStruct(Structure):
    _fields_ = [
        ("n_nodes", c_int64), ("n_edges", c_int64), ("n_relations", c_int32), ("reserved", c_int32),
        ("in_ptr", c_void_p), ("in_edge", c_void_p), ("in_eid", c_void_p), ("out_ptr", c_void_p),
        ("out_rec", c_void_p), ("dist_perm", c_void_p), ("dist_items", c_void_p), ("dist_item_ptr", c_void_p),
        ("node_order", c_void_p), ("visit_meta", c_void_p), ("bwd_stream", c_void_p), ("visit_edge_ptr", c_void_p),
    ]


class GroupsStruct(Structure):
    _fields_ = [("n_groups", c_int32), ("reserved", c_int32), ("start", c_int64 * 4), ("count", c_int64 * 4)]


# name -> (restype, argtypes); must list every symbol of include/polyphemus_b200.h (tests check this)
_P = c_void_p
SIGNATURES = {
    "pb_version": (c_int, []),
    "pb_last_error": (c_char_p, []),
    "pb_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "pb_graph_workspace_bytes": (c_size_t, [c_int64]),
    "pb_graph_count": (c_int, [_P, c_int64, _P, _P, _P, _P, _P, c_size_t, _P]),
    "pb_graph_fill": (c_int, [_P, _P, _P, c_int64, c_int32, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pb_edge_attrs_encode": (c_int, [_P, _P, c_int64, _P, _P]),
    "pb_edge_attrs_decode": (c_int, [_P, c_int64, _P, c_int64, c_int64, _P, _P, _P]),
    "pb_csr_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "pb_csr_num_dist_items": (c_int32, []),
    "pb_csr_build": (c_int, [_P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "pb_edge_table_fwd": (c_int, [_P, _P, c_int32, _P, _P]),
    "pb_edge_table_bwd": (c_int, [_P, _P, c_int32, _P, _P, _P]),
    "pb_dropout_bits_bytes": (c_size_t, [c_int64, c_int32]),
    "pb_dropout_bits": (c_int, [c_int64, c_int32, c_float, c_uint64, _P, _P]),
    "pb_agg_fwd": (c_int, [POINTER(CsrStruct), _P, c_int32, _P, _P, _P, c_int64, c_int32, _P, c_float, c_int32, _P]),
    "pb_agg_bwd": (c_int, [POINTER(CsrStruct), _P, c_int32, _P, _P, c_int64, c_int32, _P, _P, _P, _P, _P, c_float, c_int32,
                           _P]),
    "pb_dropout_mask": (c_int, [c_int64, c_int32, c_float, c_uint64, _P, _P]),
    "pb_csr_visit_meta": (c_int, [POINTER(CsrStruct), _P, _P]),
    "pb_csr_bwd_stream": (c_int, [POINTER(CsrStruct), _P, _P, _P]),
    "pb_agg_bwd_num_partials": (c_int32, [c_int64, c_int32, c_int32]),
    "pb_agg_bwd_fused": (c_int, [POINTER(CsrStruct), _P, c_int32, _P, _P, c_int64, c_int32, _P, _P, _P, _P, c_float, c_int32, _P]),
    "pb_edge_table_bwd_fused": (c_int, [_P, c_int32, c_int32, _P, _P, _P]),
    "pb_weight_prep": (c_int, [_P, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P]),
    "pb_rgcn_gemm_fwd": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32,
                                 POINTER(GroupsStruct), c_int32, c_int32, _P]),
    "pb_rgcn_gemm_fwd_bn_partial_rows": (c_int64, [c_int64]),
    "pb_rgcn_gemm_fwd_bn": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32,
                                    POINTER(GroupsStruct), c_int32, c_int32, _P, _P]),
    "pb_bn_finalize": (c_int, [_P, c_int64, c_int64, c_int32, _P, _P, c_float, c_float, _P, _P, _P, _P, _P]),
    "pb_gemm_nt": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, _P]),
    "pb_rgcn_gemm_bwd_data": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int64, c_int64, c_int32, c_int32,
                                      POINTER(GroupsStruct), c_int32, _P]),
    "pb_rgcn_gemm_bwd_weight_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "pb_rgcn_gemm_bwd_weight_workspace_bytes_for": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "pb_rgcn_gemm_bwd_weight": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_int64, c_int32, c_int32,
                                        POINTER(GroupsStruct), c_int32, _P, c_size_t, _P]),
    "pb_gemm_f32_check": (c_int, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int32, c_int32, c_int32,
                                  c_int32, _P]),
    "pb_bn_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "pb_bn_stats": (c_int, [_P, c_int64, c_int64, c_int32, POINTER(GroupsStruct), _P, _P, c_float, c_float, _P, _P, _P, _P,
                            _P, c_size_t, c_int32, _P]),
    "pb_bn_prepare_eval": (c_int, [_P, _P, _P, _P, c_float, c_int32, _P, _P]),
    "pb_bn_relu_res_fwd": (c_int, [_P, c_int64, _P, _P, _P, c_int64, c_int32, POINTER(GroupsStruct), c_int32, c_int32, _P]),
    "pb_bn_relu_res_bwd": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int64, c_int32, POINTER(GroupsStruct), c_int32, _P, _P,
                                   c_int64, _P, _P, _P, _P, c_size_t, c_int32, _P]),
    "pb_grad_prep": (c_int, [_P, c_int64, c_int64, c_int32, c_int32, _P, _P, c_int64, _P, _P, c_size_t, _P]),
    "pb_chord_embed_fwd": (c_int, [_P, c_int64, c_int32, c_int32, _P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P, c_int64,
                                   c_int64, _P]),
    "pb_chord_embed_bwd_prep": (c_int, [_P, c_int64, c_int32, c_int32, _P, c_int32, c_int32, c_int32, _P, c_int64, _P, c_int64,
                                        c_int32, _P, _P, _P, c_int64, _P]),
    "pb_bar_pool_fwd": (c_int, [_P, c_int64, _P, _P, c_int64, c_int32, _P, _P, _P]),
    "pb_bar_pool_bwd": (c_int, [_P, c_int64, _P, _P, c_int64, c_int32, _P, _P, c_int64, _P, _P]),
    "pb_bar_expand_fwd": (c_int, [_P, _P, c_int64, c_int32, _P, c_int64, _P]),
    "pb_bar_expand_bwd": (c_int, [_P, c_int64, _P, c_int64, c_int32, _P, _P]),
    "pb_ce_fwd": (c_int, [_P, c_int64, c_int32, c_int64, c_int32, _P, c_int32, _P, _P, _P]),
    "pb_ce_bwd": (c_int, [_P, c_int64, c_int32, c_int64, c_int32, _P, c_int32, _P, _P, _P, c_int64, _P]),
    "pb_ce_rows_fwd": (c_int, [_P, c_int64, c_int32, c_int64, c_int32, _P, _P, _P, _P, _P, _P]),
    "pb_ce_rows_bwd": (c_int, [_P, c_int64, c_int32, c_int64, c_int32, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "pb_rows_scatter": (c_int, [_P, c_int32, _P, c_int64, c_int32, _P, c_int32, c_int64, POINTER(GroupsStruct), _P]),
    "pb_rows_gather": (c_int, [_P, c_int32, _P, c_int64, c_int32, _P, c_int32, _P]),
    "pb_token_hist": (c_int, [_P, c_int64, c_int32, c_int32, _P, c_int64, c_int32, c_int32, _P, _P]),
    "pb_dataset_structure": (c_int, [_P, c_int64, c_int32, _P, _P]),
    "pb_dataset_tokens": (c_int, [_P, _P, _P, c_int64, c_int32, _P, _P]),
    "pb_mtp_from_logits": (c_int, [_P, c_int64, c_int32, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, _P, _P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PolyphemusB200Error(
                f"{LIB_PATH} not found: build the CUDA library first (python -m polyphemus_b200.build). "
                "polyphemus_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


# ---------------------------------------------------------------------------------------------------------
# Launch accounting + optional per-call CUDA-event timing (bench.py's roofline measurement). Every entry into
# the library goes through `call`, so `launch_counter` is the number of OUR kernels launched (memsets and
# host-only queries are not counted).
LAUNCHES = {
    "pb_graph_count": 8, "pb_graph_fill": 1, "pb_edge_attrs_encode": 1, "pb_edge_attrs_decode": 1, "pb_csr_build": 13,
    "pb_edge_table_fwd": 1, "pb_edge_table_bwd": 2, "pb_edge_table_bwd_fused": 2, "pb_agg_bwd_fused": 1, "pb_csr_visit_meta": 1, "pb_rows_scatter": 2, "pb_rows_gather": 1, "pb_csr_bwd_stream": 1, "pb_agg_fwd": 1, "pb_agg_bwd": 2, "pb_dropout_mask": 1, "pb_dropout_bits": 1,
    "pb_weight_prep": 1, "pb_rgcn_gemm_fwd": 1, "pb_rgcn_gemm_fwd_bn": 1, "pb_bn_finalize": 1, "pb_rgcn_gemm_bwd_data": 1, "pb_gemm_nt": 1, "pb_rgcn_gemm_bwd_weight": 2,
    "pb_gemm_f32_check": 1, "pb_bn_stats": 2, "pb_bn_prepare_eval": 1, "pb_bn_relu_res_fwd": 1,
    "pb_bn_relu_res_bwd": 4, "pb_grad_prep": 2, "pb_ce_fwd": 1, "pb_ce_bwd": 1, "pb_ce_rows_fwd": 1, "pb_ce_rows_bwd": 1, "pb_chord_embed_fwd": 1, "pb_chord_embed_bwd_prep": 1,
    "pb_bar_pool_fwd": 1, "pb_bar_pool_bwd": 1, "pb_bar_expand_fwd": 1, "pb_bar_expand_bwd": 1,
}
launch_counter = {"n": 0}


class EventProfiler:
    """Brackets every library call with CUDA events on the launching stream; resolved after a synchronize."""

    def __init__(self):
        self.records = []       # (name, start_event, end_event, meta)
        self.enabled = False
        self._pool = []         # timing events are reused across resets: creating them costs more than recording them
        self._used = 0

    def reset(self):
        self.records = []
        self._used = 0

    def reserve(self, n_calls: int) -> None:
        """Create the events of `n_calls` bracketed calls ahead of a timed region."""
        need = self._used + 2 * n_calls - len(self._pool)
        if need > 0:
            self._pool.extend(torch.cuda.Event(enable_timing=True) for _ in range(need))

    def events(self):
        """Two timing events from the pool (grown on demand)."""
        if self._used + 2 > len(self._pool):
            self._pool.extend(torch.cuda.Event(enable_timing=True) for _ in range(1024))
        e0, e1 = self._pool[self._used], self._pool[self._used + 1]
        self._used += 2
        return e0, e1

    def summary(self):
        """{name: {"calls": n, "ms": total_ms, "meta": last meta}} — call after torch.cuda.synchronize()."""
        out = {}
        for name, e0, e1, meta in self.records:
            ent = out.setdefault(name, {"calls": 0, "ms": 0.0, "meta": meta})
            ent["calls"] += 1
            ent["ms"] += e0.elapsed_time(e1)
        return out


profiler = EventProfiler()


def call(name: str, *args, tag=None) -> None:
    """Invoke an ABI entry point. `tag` only labels the profiler record (e.g. "linear" for the dense layers next to
    the path, so that they are not counted as message-passing launches)."""
    fn = getattr(lib(), name)
    if profiler.enabled:
        e0, e1 = profiler.events()
        e0.record()
        rc = fn(*args)
        e1.record()
        profiler.records.append((name if tag is None else f"{name}:{tag}", e0, e1, None))
    else:
        rc = fn(*args)
    check(rc, name)
    launch_counter["n"] += LAUNCHES.get(name, 1)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pb_last_error()
        raise PolyphemusB200Error(f"{what} failed with status {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    """Raw handle of the current stream of the current device (the callers sit inside ``on_device``). The private
    accessor costs ~0.3 us against ~15 us for torch.cuda.current_stream(): at ~300 library calls per training step
    that is a millisecond of host time per step."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(dev):
    """``with on_device(t.device):`` — torch.cuda.device(dev) only when `dev` is not already current (entering and
    leaving the real guard costs ~10 us, and every op of the path used to do it)."""
    if dev.index is None or dev.index == torch._C._cuda_getDevice():
        return _NO_GUARD
    return torch.cuda.device(dev)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PolyphemusB200Error(
                "polyphemus_b200 kernels only run on CUDA tensors (sm_100a); there is no CPU fallback")
