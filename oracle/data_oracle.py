"""numpy restatement of the reference's dataset-sample decode and of its logits -> multitrack-pianoroll scatter.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``dataset_item``     <- ``PolyphemusDataset.__getitem__`` data.py:218-271 on the on-disk sample layout written by
                          preprocess.py:210 (``c_tensor`` int16 [4, T, 16, 2] = (pitch id, duration id) per track /
                          timestep / token slot, ``s_tensor`` bool [4, T], T = n_bars * 32): reshape to bars
                          (data.py:226-231), one-hot (data.py:233-259), ``graph_from_tensor`` incl. the in-place fake
                          activation of empty bars (data.py:262, 152-153), silence filter (data.py:264-266).
* ``mtp_from_logits``  <- utils.py:59-79: active cells take the node's logits in order, silent cells the silence
                          pattern (slot 0 -> pitch EOS, slots 1.. -> pitch PAD).

Pinned against the reference's own code in tests/test_oracle_cpu.py (container) and through
tests/golden/dataset_items.npz / mtp_from_logits.npz (everywhere).
"""
from __future__ import annotations

import numpy as np

from . import graph_oracle as go

N_PITCH = 131          # constants.py:28
N_DUR = 99             # constants.py:40
PITCH_EOS, PITCH_PAD = 129, 130     # constants.py:22-25


def dataset_item(c_disk: np.ndarray, s_disk: np.ndarray, n_bars: int):
    """One sample. Returns (s_tensor bool [n_bars, 4, 32] with the fake activations, tokens int64 [N, 16, 2] of the
    active cells in node order (bar, track, timestep), graph arrays of the sequence)."""
    c = np.asarray(c_disk).astype(np.int64)
    s = np.asarray(s_disk).astype(bool)
    n_tracks = c.shape[0]
    c = c.reshape(n_tracks, n_bars, -1, c.shape[2], c.shape[3]).transpose(1, 0, 2, 3, 4)     # data.py:226-229
    s = s.reshape(n_tracks, n_bars, -1).transpose(1, 0, 2).copy()                            # data.py:230-231
    arrays = go.sequence_graph(s)                  # empty bars get s[bar, 0, 0] = True in place (data.py:152-153)
    s = arrays.s_tensor
    tokens = c.reshape(-1, c.shape[-2], c.shape[-1])[s.reshape(-1)]                          # data.py:264-266
    return s, tokens, arrays


def onehot(tokens: np.ndarray) -> np.ndarray:
    """int [N, 16, 2] -> float32 [N, 16, 230] (data.py:233-259)."""
    out = np.zeros(tokens.shape[:2] + (N_PITCH + N_DUR,), dtype=np.float32)
    n, t = np.meshgrid(np.arange(tokens.shape[0]), np.arange(tokens.shape[1]), indexing="ij")
    out[n, t, tokens[..., 0]] = 1.0
    out[n, t, N_PITCH + tokens[..., 1]] = 1.0
    return out


def mtp_from_logits(c_logits: np.ndarray, s_tensor: np.ndarray) -> np.ndarray:
    """c_logits [N, n_tok, d_token], s_tensor bool [B, n_bars, 4, 32] -> [B, n_bars, 4, 32, n_tok, d_token]."""
    s = np.asarray(s_tensor).astype(bool)
    n_tok, d_tok = c_logits.shape[-2:]
    silence = np.zeros((n_tok, d_tok), dtype=c_logits.dtype)
    silence[0, PITCH_EOS] = 1.0
    silence[1:, PITCH_PAD] = 1.0
    mtp = np.empty((s.size, n_tok, d_tok), dtype=c_logits.dtype)
    mtp[s.reshape(-1)] = c_logits
    mtp[~s.reshape(-1)] = silence
    return mtp.reshape(s.shape + (n_tok, d_tok))
