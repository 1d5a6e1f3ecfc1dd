"""polyphemus_b200 — B200-native (sm_100a) message-passing hot path of the Polyphemus graph VAE.

Public surface (mirrors the reference's model.py / data.py for this path):

    graph_from_tensor, graphs_from_tensor      data.py:141 / model.py:596-607   (CUDA graph builder)
    decode_samples                             data.py:218-271                  (on-disk samples -> batched graph)
    mtp_from_logits                            utils.py:59-79                   (logits -> multitrack pianoroll)
    GCL, GCN                                   model.py:41-135, 167-208         (CUDA message passing)
    VAE, Encoder, Decoder                      model.py:448-678                 (host modules calling the path)

All arithmetic of the hot path runs in libpolyphemus_b200.so (include/polyphemus_b200.h). There is no CPU,
Triton or PyTorch fallback: importing is cheap, but any call without the built library or a CUDA device raises.
"""
from . import _ffi
from ._ffi import PolyphemusB200Error
from .conv import GCL, GCN, BatchNorm
from .data import decode_samples, mtp_from_logits
from .graph import CsrPlan, Graph, decode_edge_attrs, graph_from_tensor, graphs_from_tensor
from .ops import get_precision, launch_counter, set_bf16_activations, set_precision
from .vae import VAE, ContentDecoder, ContentEncoder, Decoder, Encoder, StructureDecoder, StructureEncoder

__version__ = "0.1.0"

__all__ = [
    "GCL", "GCN", "BatchNorm", "VAE", "Encoder", "Decoder", "ContentEncoder", "ContentDecoder", "StructureEncoder",
    "StructureDecoder", "Graph", "CsrPlan", "graph_from_tensor", "graphs_from_tensor", "decode_edge_attrs",
    "decode_samples", "mtp_from_logits", "set_precision", "get_precision", "set_bf16_activations", "launch_counter", "PolyphemusB200Error",
]
