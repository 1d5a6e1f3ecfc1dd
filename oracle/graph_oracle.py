"""numpy restatement of the reference's pianoroll-structure -> typed graph construction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Follows, function by function:

* ``bar_edges``              <- ``get_node_labels`` data.py:14-21, ``get_track_edges`` data.py:24-51,
                                ``get_onset_edges`` data.py:54-80, ``get_next_edges`` data.py:83-121
* ``sequence_graph``         <- ``graph_from_tensor`` data.py:141-204 (fake activation data.py:152-153,
                                concat order data.py:159-167, fake self-edge data.py:173-176, edge_attrs
                                data.py:179-182, node_features/is_drum/num_nodes data.py:184-186, PyG
                                ``collate(increment=True, add_batch=True)`` data.py:193-202)
* ``batch_graph``            <- PyG ``Batch.from_data_list`` as used at model.py:604 / train.py:152-156

Pinned against the reference's own code in tests/test_oracle_cpu.py (container) and through
tests/golden/graph_*.npz (everywhere). Integer results must match bit for bit.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

N_TRACKS = 4          # constants.py:4
N_TIMESTEPS = 32      # resolution 8 * 4 (model.py:437-438)
REL_TRACK = 0         # constants.py:52-58  (track t -> relation t)
REL_ONSET = N_TRACKS
REL_NEXT = N_TRACKS + 1
N_RELATIONS = N_TRACKS + 2


def bar_edges(bar: np.ndarray) -> tuple[np.ndarray, int]:
    """bar: bool [4, 32] (already containing the fake activation if it was empty).

    Returns (edges int64 [E, 4] rows (u, v, type, dist) in reference order, n_nodes).
    For an edgeless bar returns the fake self-edge row (0, 0, 0, 0).
    """
    bar = np.asarray(bar).astype(bool)
    n_tracks, n_ts = bar.shape
    # node label = rank in row-major nonzero order (data.py:14-21)
    label = np.cumsum(bar.reshape(-1)).reshape(bar.shape) - 1
    n_nodes = int(bar.sum())
    rows: list[tuple[int, int, int, int]] = []

    # TRACK edges, data.py:36-49: per track, consecutive active timesteps; forward list then inverse list
    for trk in range(n_tracks):
        ts = np.flatnonzero(bar[trk])
        fwd = [(int(label[trk, a]), int(label[trk, b]), REL_TRACK + trk, int(b - a))
               for a, b in zip(ts[:-1], ts[1:])]
        rows += fwd + [(v, u, t, d) for (u, v, t, d) in fwd]

    # ONSET edges, data.py:67-78: per timestep, combinations of active tracks; forward then inverse
    for t in range(n_ts):
        trks = np.flatnonzero(bar[:, t])
        fwd = [(int(label[a, t]), int(label[b, t]), REL_ONSET, 0)
               for a, b in itertools.combinations(trks.tolist(), 2)]
        rows += fwd + [(v, u, ty, d) for (u, v, ty, d) in fwd]

    # NEXT edges, data.py:95-119: consecutive *active* timesteps, cross-track product, forward only
    active_ts = np.flatnonzero(bar.any(axis=0))
    if active_ts.size >= 2:   # a single active timestep squeezes to 0-d -> no edges (data.py:96-98)
        for t1, t2 in zip(active_ts[:-1], active_ts[1:]):
            for a in np.flatnonzero(bar[:, t1]):
                for b in np.flatnonzero(bar[:, t2]):
                    if a != b:
                        rows.append((int(label[a, t1]), int(label[b, t2]), REL_NEXT, int(t2 - t1)))

    if not rows:              # data.py:173-176
        rows = [(0, 0, 0, 0)]
    return np.asarray(rows, dtype=np.int64).reshape(-1, 4), n_nodes


@dataclass
class GraphArrays:
    edge_index: np.ndarray      # int64 [2, E]
    edge_type: np.ndarray       # int64 [E]   (edge_attrs[:, 0])
    edge_dist: np.ndarray       # int64 [E]   (argmax of edge_attrs[:, 1:])
    node_features: np.ndarray   # float32 [N, 4]
    is_drum: np.ndarray         # bool [N]
    bars: np.ndarray            # int64 [N]   bar index inside its sequence
    batch: np.ndarray           # int64 [N]   sequence index (all zeros for a single sequence)
    num_nodes: int
    s_tensor: np.ndarray        # bool, same shape as the input, with fake activations applied

    @property
    def edge_attrs(self) -> np.ndarray:
        """float32 [E, 33]: col 0 = type, col 1+dist = 1 (data.py:179-182)."""
        e = self.edge_type.shape[0]
        out = np.zeros((e, N_TIMESTEPS + 1), dtype=np.float32)
        out[:, 0] = self.edge_type
        out[np.arange(e), self.edge_dist + 1] = 1.0
        return out


def sequence_graph(s_tensor: np.ndarray) -> GraphArrays:
    """s_tensor: bool [n_bars, 4, 32]. Restates ``graph_from_tensor`` (data.py:141-204)."""
    s = np.array(s_tensor, dtype=bool, copy=True)
    ei, et, ed, nf, bars = [], [], [], [], []
    offset = 0
    for b in range(s.shape[0]):
        if not s[b].any():
            s[b, 0, 0] = True                      # data.py:152-153 (in place on the caller's tensor)
        edges, n = bar_edges(s[b])
        ei.append(edges[:, :2].T + offset)         # collate(increment=True)
        et.append(edges[:, 2])
        ed.append(edges[:, 3])
        trk = np.nonzero(s[b])[0]                  # data.py:124-138
        feat = np.zeros((n, N_TRACKS), dtype=np.float32)
        feat[np.arange(n), trk] = 1.0
        nf.append(feat)
        bars.append(np.full(n, b, dtype=np.int64))
        offset += n
    node_features = np.concatenate(nf, axis=0)
    return GraphArrays(
        edge_index=np.concatenate(ei, axis=1).astype(np.int64),
        edge_type=np.concatenate(et).astype(np.int64),
        edge_dist=np.concatenate(ed).astype(np.int64),
        node_features=node_features,
        is_drum=node_features[:, 0].astype(bool),
        bars=np.concatenate(bars),
        batch=np.zeros(offset, dtype=np.int64),
        num_nodes=offset,
        s_tensor=s,
    )


def batch_graph(s_tensor: np.ndarray) -> GraphArrays:
    """s_tensor: bool [B, n_bars, 4, 32]. Per-sequence graphs collated as Batch.from_data_list does."""
    seqs = [sequence_graph(s) for s in np.asarray(s_tensor)]
    offs = np.cumsum([0] + [g.num_nodes for g in seqs])
    return GraphArrays(
        edge_index=np.concatenate([g.edge_index + o for g, o in zip(seqs, offs)], axis=1),
        edge_type=np.concatenate([g.edge_type for g in seqs]),
        edge_dist=np.concatenate([g.edge_dist for g in seqs]),
        node_features=np.concatenate([g.node_features for g in seqs], axis=0),
        is_drum=np.concatenate([g.is_drum for g in seqs]),
        bars=np.concatenate([g.bars for g in seqs]),
        batch=np.concatenate([np.full(g.num_nodes, i, dtype=np.int64) for i, g in enumerate(seqs)]),
        num_nodes=int(offs[-1]),
        s_tensor=np.stack([g.s_tensor for g in seqs]),
    )


def synthetic_structure(batch: int, n_bars: int, p: float = 0.25, seed: int = 0) -> np.ndarray:
    """Bernoulli(p) structure tensor bool [B, n_bars, 4, 32] (SURVEY.md §8d synthetic inputs)."""
    rng = np.random.default_rng(seed)
    return rng.random((batch, n_bars, N_TRACKS, N_TIMESTEPS)) < p
