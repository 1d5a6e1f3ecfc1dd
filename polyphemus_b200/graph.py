"""Batched on-device graph construction and CSR plans.

Host-side mirror of the reference's ``graph_from_tensor`` (data.py:141-204) and of the per-sequence loop +
``Batch.from_data_list`` in ``Decoder._structure_from_binary`` (model.py:596-607): same attribute names
(``edge_index, edge_attrs, node_features, is_drum, bars, batch, num_nodes``), same values bit for bit, same
in-place fake activation of empty bars — but the whole batch is built by two CUDA kernels
(pb_graph_count / pb_graph_fill) instead of Python loops, and ``edge_attrs`` (E x 33 fp32) is only
materialised if somebody asks for it.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _ffi

N_TRACKS = 4
N_TIMESTEPS = 32


class CsrPlan:
    """Destination-/source-sorted views of an edge list (pb_csr_build); shared by all layers of a GCN."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, edge_dist: torch.Tensor,
                 num_nodes: int, n_relations: int = _ffi.N_RELATIONS, node_order: Optional[torch.Tensor] = None):
        """``node_order`` (int32 permutation of the rows): visiting order of the kernels, see ``set_node_order``; given here
        the visit metadata and the backward's record stream are built once instead of once per order."""
        _ffi.require_cuda(edge_index, edge_type, edge_dist)
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be int64 [2, E]")
        edge_index = edge_index.contiguous()
        edge_type = edge_type.contiguous()
        edge_dist = edge_dist.contiguous()
        if edge_type.dtype != torch.uint8 or edge_dist.dtype != torch.uint8:
            raise ValueError("edge_type / edge_dist must be uint8")
        dev = edge_index.device
        n, e, r = int(num_nodes), int(edge_index.size(1)), int(n_relations)
        lib = _ffi.lib()
        self.n_nodes, self.n_edges, self.n_relations = n, e, r
        self.in_ptr = torch.empty(n * r + 1, dtype=torch.int32, device=dev)
        self.in_edge = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        self.in_eid = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        self.out_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.out_rec = torch.empty((max(e, 1), 4), dtype=torch.int32, device=dev)
        self.n_dist_items = int(lib.pb_csr_num_dist_items())
        self.dist_perm = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        self.dist_items = torch.empty((self.n_dist_items, 4), dtype=torch.int32, device=dev)
        self.dist_item_ptr = torch.empty(_ffi.N_DISTS + 1, dtype=torch.int32, device=dev)
        ws_bytes = lib.pb_csr_workspace_bytes(n, e, r)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _ffi.on_device(dev):
            _ffi.call("pb_csr_build", edge_index.data_ptr(), edge_type.data_ptr(), edge_dist.data_ptr(), n, e, r,
                      self.in_ptr.data_ptr(), self.in_edge.data_ptr(), self.in_eid.data_ptr(),
                      self.out_ptr.data_ptr(), self.out_rec.data_ptr(), self.dist_perm.data_ptr(),
                      self.dist_items.data_ptr(), self.dist_item_ptr.data_ptr(), ws.data_ptr(), ws_bytes, _ffi.stream())
        self.device = dev
        self.struct = _ffi.CsrStruct(n, e, r, 0, self.in_ptr.data_ptr(), self.in_edge.data_ptr(),
                                     self.in_eid.data_ptr(), self.out_ptr.data_ptr(), self.out_rec.data_ptr(),
                                     self.dist_perm.data_ptr(), self.dist_items.data_ptr(),
                                     self.dist_item_ptr.data_ptr(), None, None, None, None)
        self.node_order = None
        if node_order is not None:
            assert node_order.dtype == torch.int32 and node_order.numel() == n
            self.node_order = node_order.contiguous()
            self.struct.node_order = self.node_order.data_ptr()
        self.visit_meta = torch.empty((n, 4), dtype=torch.int32, device=dev)
        self._build_visit_meta()

    def _build_visit_meta(self) -> None:
        """Per visited node {row, first out-edge, out-degree, first in-edge} (pb_csr_visit_meta) and the fused
        backward's flat record stream in visiting order (pb_csr_bwd_stream)."""
        st = self.struct
        st.visit_meta = st.bwd_stream = st.visit_edge_ptr = None
        with _ffi.on_device(self.device):
            _ffi.call("pb_csr_visit_meta", ctypes.byref(st), self.visit_meta.data_ptr(), _ffi.stream())
            st.visit_meta = self.visit_meta.data_ptr()
            self.visit_edge_ptr = torch.zeros(self.n_nodes + 1, dtype=torch.int32, device=self.device)
            torch.cumsum(self.visit_meta[:, 2], 0, dtype=torch.int32, out=self.visit_edge_ptr[1:])
            self.bwd_stream = torch.empty((3 * self.n_nodes + self.n_edges, 4), dtype=torch.int32, device=self.device)
            _ffi.call("pb_csr_bwd_stream", ctypes.byref(st), self.visit_edge_ptr.data_ptr(), self.bwd_stream.data_ptr(),
                      _ffi.stream())
        st.bwd_stream, st.visit_edge_ptr = self.bwd_stream.data_ptr(), self.visit_edge_ptr.data_ptr()

    def set_node_order(self, order: torch.Tensor) -> None:
        """Visit the rows in `order` (int32 permutation of 0..n_nodes-1) instead of 0..n-1 (see pb_csr_t)."""
        assert order.dtype == torch.int32 and order.numel() == self.n_nodes
        self.node_order = order.contiguous()
        self.struct.node_order = self.node_order.data_ptr()
        self._build_visit_meta()

    def ref(self):
        return ctypes.byref(self.struct)


class StructuredPlan:
    """Track-relation-sorted, 128-row-padded node layout + its CSR plan (3 slots: track / onset / next).

    A node only ever receives TRACK edges of ONE relation (its own track; relation 0 for the lone node of a
    one-node bar, whose only in-edge is the fake self-edge of data.py:173-176). Sorting the nodes by that relation
    and padding every group to a multiple of the GEMM's 128-row tile lets each tile contract
    [H_track | H_onset | H_next | x] (4d wide) against [weight[g]; weight[4]; weight[5]; root] instead of the
    7d-wide operand with three structurally-zero blocks: identical results, 4/7 of the flops and operand bytes.
    """

    TILE = 128

    def __init__(self, graph: "Graph"):
        dev = graph.edge_index.device
        counts = [int(c) for c in graph.group_counts]
        n = graph.num_nodes
        assert sum(counts) == n
        padded = [(c + self.TILE - 1) // self.TILE * self.TILE for c in counts]
        starts = [sum(padded[:g]) for g in range(4)]
        cums = [sum(counts[:g]) for g in range(4)]
        self.n_padded = max(sum(padded), self.TILE)
        self.counts, self.starts = counts, starts
        group = graph.node_group.to(torch.int16)
        order = torch.argsort(group, stable=True)                       # nodes in group order (original order kept)
        shift = torch.tensor([starts[g] - cums[g] for g in range(4)], dtype=torch.int64, device=dev)
        pos_sorted = torch.arange(n, device=dev) + shift.index_select(0, group.index_select(0, order).long())
        self.pos = torch.empty(n, dtype=torch.int64, device=dev).index_copy_(0, order, pos_sorted)  # node -> padded row
        slot = (graph.edge_type.to(torch.int16) - 3).clamp_(min=0).to(torch.uint8)   # track rels -> 0, onset 1, next 2
        # visit the rows bar by bar (original node order), the padding rows last: a bar's rows live in four distant
        # group regions, touching them together keeps every gathered row in L2 until its last use
        is_pad = torch.ones(self.n_padded, dtype=torch.bool, device=dev).index_fill_(0, self.pos, False)
        pads = torch.arange(self.n_padded, device=dev)[is_pad] if self.n_padded > n else self.pos[:0]
        order = torch.cat((self.pos, pads)).to(torch.int32)
        self.plan = CsrPlan(self.pos[graph.edge_index], slot, graph.edge_dist, self.n_padded, n_relations=3,
                            node_order=order)
        self.groups = _ffi.GroupsStruct(4, 0, (ctypes.c_int64 * 4)(*starts), (ctypes.c_int64 * 4)(*counts))

    def groups_ref(self):
        return ctypes.byref(self.groups)


class Graph:
    """Attribute bag with the fields of the reference's PyG ``Data``/``Batch`` for this path."""

    def __init__(self, **kwargs):
        self._edge_attrs = None
        self._plan = None
        self._structured = None
        for k, v in kwargs.items():
            setattr(self, k, v)

    # edge_attrs f32 [E, 33] (data.py:179-182) is derived data: built on first access
    @property
    def edge_attrs(self) -> torch.Tensor:
        if self._edge_attrs is None:
            e = self.edge_type.numel()
            out = torch.empty((e, N_TIMESTEPS + 1), dtype=torch.float32, device=self.edge_type.device)
            with _ffi.on_device(out.device):
                _ffi.call("pb_edge_attrs_encode", self.edge_type.data_ptr(), self.edge_dist.data_ptr(), e,
                          out.data_ptr(), _ffi.stream())
            self._edge_attrs = out
        return self._edge_attrs

    @edge_attrs.setter
    def edge_attrs(self, value):
        """Replacing the edge attributes replaces the graph's relations / distances: everything derived from them
        (uint8 type / distance, the CSR plans, the structured layout — only valid for builder graphs) is dropped."""
        self._edge_attrs = value
        self._plan = None
        self._structured = None
        if value is not None and "edge_type" in self.__dict__:
            self.edge_type, self.edge_dist = decode_edge_attrs(value[:, 0], value[:, 1:])
            self.group_counts = None          # no longer known to be a builder graph

    @property
    def plan(self) -> CsrPlan:
        if self._plan is None:
            self._plan = CsrPlan(self.edge_index, self.edge_type, self.edge_dist, self.num_nodes)
        return self._plan

    @property
    def structured(self) -> Optional[StructuredPlan]:
        """Structured layout, available for graphs that came out of the device builder."""
        if getattr(self, "_structured", None) is None:
            if getattr(self, "group_counts", None) is None or getattr(self, "node_group", None) is None:
                return None
            self._structured = StructuredPlan(self)
        return self._structured

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")] + ["edge_attrs"]

    def to(self, device, *args, **kwargs):
        device = torch.device(device)
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor) and v.device != device:
                if k.startswith("_"):
                    setattr(self, k, None)
                else:
                    setattr(self, k, v.to(device, *args, **kwargs))
        if self._plan is not None and self._plan.device != device:
            self._plan = None
        if self._structured is not None and self._structured.pos.device != device:
            self._structured = None
        return self


def _as_device_bytes(s_tensor: torch.Tensor, device: Optional[torch.device]):
    """Returns (uint8 view on CUDA sharing storage when possible, write-back target or None)."""
    if s_tensor.dtype not in (torch.bool, torch.uint8):
        raise ValueError("s_tensor must be a bool/uint8 structure tensor")
    if s_tensor.size(-1) != N_TIMESTEPS or s_tensor.size(-2) != N_TRACKS:
        raise ValueError(f"s_tensor must end in [{N_TRACKS}, {N_TIMESTEPS}], got {tuple(s_tensor.shape)}")
    if s_tensor.is_cuda and s_tensor.is_contiguous():
        return s_tensor.view(torch.uint8), None
    if device is None:
        device = s_tensor.device if s_tensor.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise _ffi.PolyphemusB200Error("graph construction runs on the GPU only (no CPU fallback)")
    dev_copy = s_tensor.to(device).contiguous()
    return dev_copy.view(torch.uint8), s_tensor


def graphs_from_tensor(s_tensor: torch.Tensor, device: Optional[torch.device] = None,
                       with_edge_attrs: bool = False) -> Graph:
    """s_tensor bool [B, n_bars, 4, 32] -> one batched graph (== Batch.from_data_list of per-sequence graphs).

    Empty bars get the fake activation ``s[b, bar, 0, 0] = True`` in the caller's tensor, exactly as
    data.py:152-153 does. One host sync (the {N, E} read-back that sizes the outputs).
    """
    if s_tensor.dim() != 4:
        raise ValueError("expected s_tensor of shape [B, n_bars, 4, 32]")
    bsz, n_bars = int(s_tensor.size(0)), int(s_tensor.size(1))
    s_u8, write_back = _as_device_bytes(s_tensor, device)
    dev = s_u8.device
    lib = _ffi.lib()
    total_bars = bsz * n_bars
    with _ffi.on_device(dev):
        st = _ffi.stream()
        bar_bits = torch.empty((total_bars, 4), dtype=torch.int32, device=dev)
        node_ptr = torch.empty(total_bars + 1, dtype=torch.int32, device=dev)
        edge_ptr = torch.empty(total_bars + 1, dtype=torch.int32, device=dev)
        totals = torch.empty(8, dtype=torch.int64, device=dev)
        ws_bytes = lib.pb_graph_workspace_bytes(total_bars)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _ffi.call("pb_graph_count", s_u8.data_ptr(), total_bars, bar_bits.data_ptr(), node_ptr.data_ptr(),
                  edge_ptr.data_ptr(), totals.data_ptr(), ws.data_ptr(), ws_bytes, st)
        n, e, n_drum, _, *group_counts = (int(v) for v in totals.tolist())          # the one host sync
        edge_index = torch.empty((2, e), dtype=torch.int64, device=dev)
        edge_type = torch.empty(e, dtype=torch.uint8, device=dev)
        edge_dist = torch.empty(e, dtype=torch.uint8, device=dev)
        edge_attrs = torch.empty((e, N_TIMESTEPS + 1), dtype=torch.float32, device=dev) if with_edge_attrs else None
        node_features = torch.empty((n, N_TRACKS), dtype=torch.float32, device=dev)
        is_drum = torch.empty(n, dtype=torch.bool, device=dev)
        bars = torch.empty(n, dtype=torch.int64, device=dev)
        batch = torch.empty(n, dtype=torch.int64, device=dev)
        node_track = torch.empty(n, dtype=torch.uint8, device=dev)
        node_group = torch.empty(n, dtype=torch.uint8, device=dev)
        _ffi.call("pb_graph_fill", bar_bits.data_ptr(), node_ptr.data_ptr(), edge_ptr.data_ptr(), total_bars, n_bars,
                  edge_index.data_ptr(), e, edge_type.data_ptr(), edge_dist.data_ptr(), _ffi.ptr(edge_attrs),
                  node_features.data_ptr(), is_drum.data_ptr(), bars.data_ptr(), batch.data_ptr(),
                  node_track.data_ptr(), node_group.data_ptr(), st)
    if write_back is not None:
        write_back.copy_(s_u8.view(write_back.dtype).reshape(write_back.shape))
    g = Graph(edge_index=edge_index, edge_type=edge_type, edge_dist=edge_dist, node_features=node_features,
              is_drum=is_drum, bars=bars, batch=batch, num_nodes=n, num_edges=e, num_graphs=bsz, n_bars=n_bars,
              n_drum=n_drum, node_track=node_track, node_group=node_group, group_counts=group_counts,
              bar_ptr=node_ptr, bar_bits=bar_bits)
    if edge_attrs is not None:
        g._edge_attrs = edge_attrs
    return g


def graph_from_tensor(s_tensor: torch.Tensor, device: Optional[torch.device] = None) -> Graph:
    """Drop-in for ``data.graph_from_tensor`` (data.py:141): one sequence ``[n_bars, 4, 32]``.

    As in the reference, ``graph.batch`` and ``graph.bars`` both hold the bar index of each node.
    """
    if s_tensor.dim() != 3:
        raise ValueError("expected s_tensor of shape [n_bars, 4, 32]")
    g = graphs_from_tensor(s_tensor.unsqueeze(0), device=device)
    g.batch = g.bars
    g.bar_ptr = None        # `batch` no longer indexes sequences: the per-bar segment operators do not apply
    return g


def plan_for(data, num_nodes: Optional[int] = None) -> CsrPlan:
    """CSR plan for whatever GCN/GCL was handed: our own Graph, or a foreign PyG-style object."""
    if isinstance(data, Graph):
        return data.plan
    cached = getattr(data, "_pb_plan", None)
    if cached is not None:
        return cached
    edge_attrs = data.edge_attrs
    edge_type, edge_dist = decode_edge_attrs(edge_attrs[:, 0], edge_attrs[:, 1:])
    n = int(num_nodes if num_nodes is not None else data.x.size(0))
    plan = CsrPlan(data.edge_index, edge_type, edge_dist, n)
    try:
        data._pb_plan = plan
    except Exception:
        pass
    return plan


def decode_edge_attrs(edge_type: torch.Tensor, edge_attr: torch.Tensor):
    """float edge_type [E] + one-hot edge_attr [E, 32] (model.py:193-194) -> uint8 type, uint8 distance."""
    _ffi.require_cuda(edge_type, edge_attr)
    if edge_attr.dim() != 2 or edge_attr.size(1) != _ffi.N_DISTS:
        raise ValueError("edge_attr must be [E, 32] one-hot timestep distances")
    edge_type = edge_type.float()
    edge_attr = edge_attr.float()
    if edge_attr.stride(1) != 1:
        edge_attr = edge_attr.contiguous()
    e = int(edge_attr.size(0))
    dev = edge_attr.device
    t_out = torch.empty(e, dtype=torch.uint8, device=dev)
    d_out = torch.empty(e, dtype=torch.uint8, device=dev)
    if e:
        with _ffi.on_device(dev):
            _ffi.call("pb_edge_attrs_decode", edge_type.data_ptr(), edge_type.stride(0), edge_attr.data_ptr(),
                      edge_attr.stride(0), e, t_out.data_ptr(), d_out.data_ptr(), _ffi.stream())
    return t_out, d_out
