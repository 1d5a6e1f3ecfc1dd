"""Relational graph-convolution modules with the reference's surface, backed by the sm_100a kernels.

Drop-in for ``GCL`` (model.py:41-135) and ``GCN`` (model.py:167-208): same constructor arguments, same
``forward`` signatures (PyG-style ``conv(x, edge_index, edge_type, edge_attr)`` / ``gcn(data)``), same
parameter names and shapes (``layers.{i}.weight/root/bias/nn.weight/nn.bias``,
``norm_layers.{i}.module.*``), same initialisation order under a given seed. The arithmetic is entirely in
libpolyphemus_b200 (see ops.py); there is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .graph import CsrPlan, Graph, decode_edge_attrs, plan_for


def _glorot_(t: torch.Tensor) -> None:
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-bound, bound)


def _reset(module: Optional[nn.Module]) -> None:
    if module is None:
        return
    kids = list(module.children())
    if kids:
        for kid in kids:
            _reset(kid)
    elif hasattr(module, "reset_parameters"):
        module.reset_parameters()


class GCL(nn.Module):
    """Edge-conditioned relational graph conv: ``out = sum_r mean_r(relu(x_j * nn(e_ij))) @ W_r + x @ root + b``."""

    def __init__(self, in_channels: int, out_channels: int, num_relations: int, nn: nn.Module,
                 dropout: float = 0.1, aggr: str = "mean", root_weight: bool = True, bias: bool = True,
                 precision: Optional[str] = None, **kwargs):
        super().__init__()
        if kwargs.get("num_bases") is not None or kwargs.get("num_blocks") is not None:
            raise NotImplementedError("basis / block-diagonal decompositions are not used by Polyphemus")
        if aggr != "mean" or not root_weight:
            raise NotImplementedError("the CUDA path implements aggr='mean' with a root weight (RGCNConv defaults)")
        self.in_channels = in_channels
        self.in_channels_l = in_channels
        self.out_channels = out_channels
        self.num_relations = num_relations
        self.num_bases = None
        self.num_blocks = None
        self.weight = torch.nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
        self.register_parameter("comp", None)
        self.root = torch.nn.Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self.nn = nn
        self.dropout = dropout
        self.precision = precision
        self.reset_edge_nn()
        self._plan_key = None
        self._plan_src = None
        self._plan = None

    def reset_parameters(self) -> None:
        _glorot_(self.weight)
        _glorot_(self.root)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def reset_edge_nn(self) -> None:
        _reset(self.nn)

    def _edge_nn_params(self):
        lin = self.nn
        if not isinstance(lin, nn.Linear) or lin.in_features != 32 or lin.bias is None:
            raise NotImplementedError("the edge network must be nn.Linear(32, d) on one-hot timestep distances")
        return lin.weight, lin.bias

    def plan_from(self, x, edge_index, edge_type, edge_attr) -> CsrPlan:
        """CSR plan of a foreign (PyG-style) edge list. The last plan is reused only for the *same tensor objects* at
        the same in-place version: the cache holds references to them, so the caching allocator cannot hand their
        addresses to a different graph while the entry is alive (the reference rebuilds its masks on every call)."""
        src = (edge_index, edge_type, edge_attr)
        key = (edge_index._version, edge_type._version, edge_attr._version, tuple(edge_index.shape), int(x.size(0)))
        cached = self._plan_src
        if cached is None or key != self._plan_key or any(a is not b for a, b in zip(cached, src)):
            t8, d8 = decode_edge_attrs(edge_type, edge_attr)
            self._plan = CsrPlan(edge_index, t8, d8, int(x.size(0)), self.num_relations)
            self._plan_key, self._plan_src = key, src
        return self._plan

    def forward(self, x, edge_index=None, edge_type=None, edge_attr=None, *, plan: Optional[CsrPlan] = None,
                bn: Optional[nn.BatchNorm1d] = None, struct=None, drawn=None):
        """``bn`` fuses BatchNorm + ReLU + residual of the enclosing GCN layer (model.py:202-206)."""
        if isinstance(x, tuple) or x is None or x.dtype == torch.long:
            raise NotImplementedError("bipartite / index-valued node inputs are not part of the Polyphemus path")
        if plan is None:
            assert edge_type is not None
            plan = self.plan_from(x, edge_index, edge_type, edge_attr)
        nn_w, nn_b = self._edge_nn_params()
        kw = {}
        if bn is not None:
            use_batch_stats = self.training or bn.running_mean is None
            kw = dict(gamma=bn.weight, beta=bn.bias, running_mean=bn.running_mean, running_var=bn.running_var,
                      eps=bn.eps, momentum=bn.momentum if bn.momentum is not None else 0.1)
            if self.training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            training = use_batch_stats
        else:
            training = self.training
        return ops.rgc_layer(x, self.weight, self.root, self.bias, nn_w, nn_b, plan, batch_norm=bn is not None,
                             training=training, p_drop=self.dropout if self.training else 0.0,
                             precision=self.precision, struct=struct, **kw,
                             **({} if drawn is None else dict(seed=drawn[0], keep_bits=drawn[1])))

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, num_relations={self.num_relations}, dropout={self.dropout}"


class BatchNorm(nn.Module):
    """Same state-dict layout as PyG's ``BatchNorm`` wrapper (keys ``module.*``), model.py:9,180,186."""

    def __init__(self, in_channels: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum)

    def reset_parameters(self) -> None:
        self.module.reset_parameters()


class GCN(nn.Module):
    """Stack of GCL layers sharing one edge network, each followed by BatchNorm, ReLU and a residual add."""

    def __init__(self, input_dim: int = 256, hidden_dim: int = 256, n_layers: int = 3, num_relations: int = 3,
                 num_dists: int = 32, batch_norm: bool = False, dropout: float = 0.1,
                 precision: Optional[str] = None):
        super().__init__()
        self.layers = nn.ModuleList()
        self.norm_layers = nn.ModuleList()
        edge_nn = nn.Linear(num_dists, input_dim)
        self.batch_norm = batch_norm
        dims = [input_dim] + [hidden_dim] * n_layers
        for i in range(n_layers):
            self.layers.append(GCL(dims[i], dims[i + 1], num_relations, edge_nn, precision=precision))
            if batch_norm:
                self.norm_layers.append(BatchNorm(hidden_dim))
        self.p = dropout

    def forward(self, data, out_perm: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32):
        """``out_perm`` (int64 permutation of the nodes) returns the rows in that order, ``out_dtype`` in that storage
        type — both folded into the structured layout's exit gather (the content decoder wants drum nodes first and
        bf16 rows for its un-embedding GEMM)."""
        x = data.x
        st = data.structured if isinstance(data, Graph) and ops.structured_enabled() else None
        if st is not None and self.batch_norm and self.p == 0 and x.size(1) % 256 == 0 and x.is_cuda:
            # structured layout: nodes sorted by track relation, groups padded to the GEMM tile (zero rows), the
            # whole stack runs on [Np, d]; one gather in, one gather out
            # bf16 mode: the stack keeps its activations (x, the pre-BatchNorm output, y and their gradients) in bf16
            # between the kernels — 16-bit storage as under the reference's fp16 autocast, fp32 arithmetic inside
            act_bf16 = (self.layers[0].precision or ops.get_precision()) == "bf16" and ops.bf16_activations_enabled()
            # GCL.message's dropout masks of all layers (model.py:133) are drawn ahead on a side stream
            p_msg = self.layers[0].dropout if self.training else 0.0
            same_p = all(layer.dropout == p_msg or not self.training for layer in self.layers)
            drawn = ops.prefetch_keep_bits(st.plan, x.size(1), p_msg, len(self.layers)) if same_p else None
            xp = ops.ScatterRowsFn.apply(x, st, torch.bfloat16 if act_bf16 else torch.float32)
            for i, layer in enumerate(self.layers):
                xp = layer(xp, plan=st.plan, bn=self.norm_layers[i].module, struct=st,
                           drawn=None if drawn is None else drawn[i])
            rows = st if out_perm is None else _PermutedRows(st, out_perm)
            return ops.GatherRowsFn.apply(xp, rows, out_dtype if act_bf16 else torch.float32)
        plan = plan_for(data, num_nodes=x.size(0))
        for i, layer in enumerate(self.layers):
            residual = x
            h = F.dropout(x, p=self.p, training=self.training) if self.p > 0 else x
            if self.batch_norm and h is residual:
                x = layer(h, plan=plan, bn=self.norm_layers[i].module)       # fully fused layer
            else:
                h = layer(h, plan=plan)
                if self.batch_norm:
                    h = self.norm_layers[i].module(h)
                x = residual + F.relu(h)
        return x if out_perm is None else x.index_select(0, out_perm)


class _PermutedRows:
    """The structured layout's node -> padded-row map composed with a node permutation (row i <- node perm[i])."""

    def __init__(self, st, perm: torch.Tensor):
        self.pos = st.pos.index_select(0, perm)
        self.n_padded, self.groups_ref = st.n_padded, st.groups_ref


__all__ = ["GCL", "GCN", "BatchNorm", "Graph"]
