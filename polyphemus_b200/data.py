"""Device-side decode of dataset samples and the generation post-processing next to the path.

* ``decode_samples``  — what ``PolyphemusDataset.__getitem__`` (data.py:218-271) + the PyG ``DataLoader`` collation do
                        for a batch, starting from the samples' ON-DISK layout (preprocess.py:210): ``c_tensor`` int16
                        ``[4, T, 16, 2]`` and ``s_tensor`` bool ``[4, T]`` per sample, ``T = n_bars * 32``. Bars-major
                        reshape, fake activation of empty bars, batched graph construction and the silence filter all
                        run on the device; the result is the batched ``Graph`` with ``s_tensor`` (float
                        ``[B * n_bars, 4, 32]``, as the reference attaches it) and ``c_tokens`` (int16 ``[N, 16, 2]``:
                        the rows of the reference's ``c_tensor`` before its one-hot expansion).
* ``mtp_from_logits`` — ``utils.mtp_from_logits`` (utils.py:59-79) as one kernel.

No CPU fallback: both raise without a CUDA device.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _ffi
from .graph import Graph, graphs_from_tensor

N_TRACKS, N_TIMESTEPS, MAX_SIMU_TOKENS = 4, 32, 16
PITCH_EOS, PITCH_PAD = 129, 130            # constants.py:22-25


def decode_samples(c_disk: torch.Tensor, s_disk: torch.Tensor, n_bars: int, device: Optional[torch.device] = None,
                   onehot: bool = False) -> Graph:
    """c_disk int16 [B, 4, T, 16, 2], s_disk bool/uint8 [B, 4, T] (host, ideally pinned, or device) -> batched Graph."""
    if c_disk.dim() == 4:
        c_disk, s_disk = c_disk.unsqueeze(0), s_disk.unsqueeze(0)
    bsz, t_len = int(s_disk.size(0)), int(s_disk.size(-1))
    if c_disk.dtype != torch.int16 or tuple(c_disk.shape) != (bsz, N_TRACKS, t_len, MAX_SIMU_TOKENS, 2):
        raise ValueError(f"c_disk must be int16 [B, 4, T, 16, 2], got {c_disk.dtype} {tuple(c_disk.shape)}")
    if s_disk.dtype not in (torch.bool, torch.uint8) or tuple(s_disk.shape) != (bsz, N_TRACKS, t_len):
        raise ValueError(f"s_disk must be bool/uint8 [B, 4, T], got {s_disk.dtype} {tuple(s_disk.shape)}")
    if t_len != n_bars * N_TIMESTEPS:
        raise ValueError(f"T = {t_len} is not n_bars * 32 = {n_bars * N_TIMESTEPS}")
    if not torch.cuda.is_available():
        raise _ffi.PolyphemusB200Error("dataset decode runs on the GPU only (no CPU fallback)")
    if device is None:
        device = c_disk.device if c_disk.is_cuda else torch.device("cuda", torch.cuda.current_device())
    c_dev = c_disk.to(device, non_blocking=True).contiguous()
    s_dev = s_disk.to(device, non_blocking=True).contiguous()
    with _ffi.on_device(device):
        st = _ffi.stream()
        s_tensor = torch.empty((bsz, n_bars, N_TRACKS, N_TIMESTEPS), dtype=torch.bool, device=device)
        _ffi.call("pb_dataset_structure", s_dev.view(torch.uint8).data_ptr(), bsz, n_bars, s_tensor.data_ptr(), st)
        graph = graphs_from_tensor(s_tensor)                      # fake activations are written into s_tensor
        tokens = torch.empty((graph.num_nodes, MAX_SIMU_TOKENS, 2), dtype=torch.int16, device=device)
        _ffi.call("pb_dataset_tokens", c_dev.data_ptr(), graph.bar_bits.data_ptr(), graph.bar_ptr.data_ptr(), bsz, n_bars,
                  tokens.data_ptr(), st)
    graph.s_tensor = s_tensor.view(-1, N_TRACKS, N_TIMESTEPS).float()
    graph.c_tokens = tokens
    graph.c_tensor = None
    if onehot:
        from .train import onehot_content
        graph.c_tensor = onehot_content(tokens)
    return graph


def mtp_from_logits(c_logits: torch.Tensor, s_tensor: torch.Tensor) -> torch.Tensor:
    """c_logits [N, n_tok, d_token] (fp32 / bf16, CUDA), s_tensor bool [B, n_bars, 4, 32] with N active cells ->
    multitrack pianoroll [B, n_bars, 4, 32, n_tok, d_token]: active cells hold their node's logits, silent cells the
    silence pattern (slot 0: pitch EOS, other slots: pitch PAD). utils.py:59-79."""
    _ffi.require_cuda(c_logits)
    if c_logits.dim() != 3 or c_logits.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("c_logits must be a CUDA float32 / bfloat16 tensor [N, n_tok, d_token]")
    dev = c_logits.device
    active = s_tensor.to(dev).bool().reshape(-1)
    n_cells = active.numel()
    incl = torch.cumsum(active, 0, dtype=torch.int32)
    node_of_cell = (incl - active.to(torch.int32)).contiguous()
    n_tok, d_tok = int(c_logits.size(1)), int(c_logits.size(2))
    c_logits = c_logits.contiguous()
    mtp = torch.empty(tuple(s_tensor.shape) + (n_tok, d_tok), dtype=c_logits.dtype, device=dev)
    s_u8 = active.to(torch.uint8)
    with _ffi.on_device(dev):
        _ffi.call("pb_mtp_from_logits", c_logits.data_ptr(), n_tok * d_tok, _ffi.PB_BF16 if c_logits.dtype == torch.bfloat16 else _ffi.PB_F32,
                  s_u8.data_ptr(), node_of_cell.data_ptr(), n_cells, n_tok, d_tok, PITCH_EOS, PITCH_PAD, mtp.data_ptr(),
                  _ffi.stream())
    return mtp


__all__ = ["decode_samples", "mtp_from_logits"]
