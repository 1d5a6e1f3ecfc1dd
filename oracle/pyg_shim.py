"""Stand-in for the nine torch-geometric==2.0.2 / torch-sparse==0.6.9 symbols the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). The reference (``/root/reference/model.py:6-12``,
``data.py:6-8``, ``train.py:9``) imports

    torch_sparse.SparseTensor, torch_sparse.masked_select_nnz
    torch_geometric.typing.{OptTensor, Adj}
    torch_geometric.nn.inits.reset
    torch_geometric.nn.norm.BatchNorm
    torch_geometric.nn.glob.GlobalAttention
    torch_geometric.nn.conv.RGCNConv
    torch_geometric.data.{Data, Batch, Dataset}
    torch_geometric.data.collate.collate
    torch_geometric.loader.DataLoader

None of these packages exist in this image and there is no network, so the semantics below are a
restatement of PyG 2.0.2's published behaviour (SURVEY.md §8c "[PyG-recall]"), each rule kept small and
independently testable (tests/test_oracle_cpu.py). ``install()`` registers the modules in ``sys.modules``
so that the reference's files import unmodified.
"""
from __future__ import annotations

import math
import sys
import types
from typing import Optional, Union

import torch
from torch import Tensor, nn


# --------------------------------------------------------------------------- torch_sparse
class SparseTensor:  # only used in isinstance() checks by the reference (model.py:31,74)
    pass


def masked_select_nnz(*_a, **_k):  # never reached with dense edge_index
    raise NotImplementedError("SparseTensor adjacency is not part of the Polyphemus path")


# --------------------------------------------------------------------------- typing
OptTensor = Optional[Tensor]
Adj = Union[Tensor, SparseTensor]


# --------------------------------------------------------------------------- nn.inits.reset
def reset(module):
    """PyG ``reset``: call reset_parameters on the children, or on the module itself if it is a leaf."""
    if module is None:
        return
    children = list(module.children()) if hasattr(module, "children") else []
    if children:
        for child in children:
            reset(child)
    elif hasattr(module, "reset_parameters"):
        module.reset_parameters()


def _glorot(t: Tensor):
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-bound, bound)


# --------------------------------------------------------------------------- MessagePassing / RGCNConv
class RGCNConv(nn.Module):
    """Parameters + ``propagate`` of PyG 2.0.2 ``RGCNConv`` (no bases / blocks), aggr='mean'."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases=None, num_blocks=None,
                 aggr="mean", root_weight=True, bias=True, **kwargs):
        super().__init__()
        if num_bases is not None or num_blocks is not None:
            raise NotImplementedError
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.aggr = aggr
        self.in_channels = in_channels
        self.in_channels_l = in_channels[0]
        self.out_channels = out_channels
        self.num_relations = num_relations
        self.num_bases = None
        self.num_blocks = None
        self.weight = nn.Parameter(torch.empty(num_relations, in_channels[0], out_channels))
        self.register_parameter("comp", None)
        if root_weight:
            self.root = nn.Parameter(torch.empty(in_channels[1], out_channels))
        else:
            self.register_parameter("root", None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        _glorot(self.weight)
        if self.root is not None:
            _glorot(self.root)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs.pop("x")
        src, dst = edge_index[0], edge_index[1]
        x_j = x.index_select(0, src)                       # __lift__, flow source_to_target
        msg = self.message(x_j=x_j, **kwargs)
        n_dst = size[1] if size is not None else x.size(0)
        out = torch.zeros(n_dst, msg.size(-1), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, dst, msg)
        if self.aggr == "mean":                            # torch_scatter.scatter_mean
            cnt = torch.zeros(n_dst, dtype=msg.dtype, device=msg.device)
            cnt.index_add_(0, dst, torch.ones_like(dst, dtype=msg.dtype))
            out = out / cnt.clamp(min=1).unsqueeze(-1)
        return out

    def message(self, x_j, **kwargs):
        return x_j


# --------------------------------------------------------------------------- nn.norm.BatchNorm
class BatchNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x):
        return self.module(x)


# --------------------------------------------------------------------------- nn.glob.GlobalAttention
class GlobalAttention(nn.Module):
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn
        reset(self.gate_nn)
        reset(self.nn)

    def forward(self, x, batch, size=None):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        size = int(batch[-1].item()) + 1 if size is None else size
        gate = self.gate_nn(x).view(-1, 1)
        x = self.nn(x) if self.nn is not None else x
        # softmax(gate, batch): subtract the segment max, exp, divide by (segment sum + 1e-16)
        seg_max = torch.full((size, 1), float("-inf"), dtype=gate.dtype, device=gate.device)
        seg_max = seg_max.scatter_reduce(0, batch.view(-1, 1), gate, reduce="amax", include_self=True)
        e = (gate - seg_max.index_select(0, batch)).exp()
        seg_sum = torch.zeros(size, 1, dtype=gate.dtype, device=gate.device).index_add_(0, batch, e)
        gate = e / (seg_sum.index_select(0, batch) + 1e-16)
        out = torch.zeros(size, x.size(-1), dtype=x.dtype, device=x.device).index_add_(0, batch, gate * x)
        return out


# --------------------------------------------------------------------------- data.Data / Batch / collate
class Data:
    """Attribute bag with the handful of PyG ``Data`` behaviours the reference relies on."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k in self.__dict__.keys() if not k.startswith("_")]

    def to(self, device, *args, **kwargs):
        for k in self.keys:
            v = getattr(self, k)
            if isinstance(v, Tensor):
                setattr(self, k, v.to(device, *args, **kwargs))
        return self

    def __cat_dim__(self, key, value):
        return -1 if ("index" in key or "face" in key) else 0

    def __inc__(self, key, value, num_nodes):
        if "batch" in key:
            return int(value.max()) + 1
        if "index" in key or "face" in key:
            return num_nodes
        return 0


def collate(cls, data_list, increment=True, add_batch=True, follow_batch=None, exclude_keys=None):
    exclude = set(exclude_keys or [])
    out = cls()
    keys = [k for k in data_list[0].keys if k not in exclude and k != "ptr"]
    num_nodes_list = [int(d.num_nodes) for d in data_list]
    for key in keys:
        if key == "num_nodes":
            continue
        vals = [getattr(d, key) for d in data_list]
        if isinstance(vals[0], Tensor):
            cat_dim = data_list[0].__cat_dim__(key, vals[0])
            if increment:
                shifted, inc = [], 0
                for d, v, n in zip(data_list, vals, num_nodes_list):
                    shifted.append(v + inc if inc != 0 else v)
                    inc += d.__inc__(key, v, n)
                vals = shifted
            if vals[0].dim() == 0:
                setattr(out, key, torch.stack(vals))
            else:
                setattr(out, key, torch.cat(vals, dim=cat_dim))
        else:
            setattr(out, key, vals)
    out.num_nodes = sum(num_nodes_list)
    if add_batch:
        counts = torch.tensor(num_nodes_list, dtype=torch.long)
        dev = None
        for key in keys:
            v = getattr(out, key, None)
            if isinstance(v, Tensor):
                dev = v.device
                break
        out.batch = torch.repeat_interleave(torch.arange(len(data_list)), counts).to(dev)
        out.ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).to(dev)
    return out, None, None


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        out, _, _ = collate(cls, data_list, increment=True, add_batch=True, exclude_keys=exclude_keys)
        out.num_graphs = len(data_list)
        return out


class Dataset(torch.utils.data.Dataset):
    def __init__(self, *a, **k):
        super().__init__()


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)


# --------------------------------------------------------------------------- registration
def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-in modules (idempotent). Also stubs ``muspy`` / ``prettytable`` for utils.py."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_pb_shim", False):
        return
    _mod("torch_sparse", SparseTensor=SparseTensor, masked_select_nnz=masked_select_nnz)
    tg = _mod("torch_geometric", _pb_shim=True)
    tg.typing = _mod("torch_geometric.typing", OptTensor=OptTensor, Adj=Adj)
    tg.nn = _mod("torch_geometric.nn")
    tg.nn.inits = _mod("torch_geometric.nn.inits", reset=reset)
    tg.nn.norm = _mod("torch_geometric.nn.norm", BatchNorm=BatchNorm)
    tg.nn.glob = _mod("torch_geometric.nn.glob", GlobalAttention=GlobalAttention)
    tg.nn.conv = _mod("torch_geometric.nn.conv", RGCNConv=RGCNConv)
    tg.data = _mod("torch_geometric.data", Data=Data, Batch=Batch, Dataset=Dataset)
    tg.data.collate = _mod("torch_geometric.data.collate", collate=collate)
    tg.loader = _mod("torch_geometric.loader", DataLoader=DataLoader)
    for absent in ("muspy", "prettytable"):
        if absent not in sys.modules:
            try:
                __import__(absent)
            except Exception:
                _mod(absent, PrettyTable=object)
