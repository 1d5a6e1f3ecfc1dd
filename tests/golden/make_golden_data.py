"""Generate tests/golden/dataset_items.npz and mtp_from_logits.npz by executing the reference's OWN unmodified
``PolyphemusDataset.__getitem__`` (data.py:218-271) and ``utils.mtp_from_logits`` (utils.py:59-79) (container only).

    python tests/golden/make_golden_data.py

Synthetic samples in the on-disk layout of preprocess.py:210 are written to a temporary directory as ``.npz`` files,
read back through the reference's dataset class, and inputs + outputs are stored. Edge cases: an empty bar (fake
activation, whose tokens are the silent cell's [SOS, EOS, PAD...]), a full bar, a single-node bar.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def samples(n_bars: int, seed: int):
    rng = np.random.default_rng(seed)
    t_len = n_bars * 32
    out = []
    for k in range(4):
        s = rng.random((4, t_len)) < (0.1, 0.25, 0.5, 0.25)[k]
        if k == 0:
            s[:, :32] = False                    # empty first bar
        if k == 1:
            s[:, 32:64] = True                   # full second bar
        if k == 2:
            s[:, -32:] = False
            s[2, -5] = True                      # single-node last bar
        c = np.empty((4, t_len, 16, 2), dtype=np.int16)
        c[..., 0], c[..., 1] = 130, 98
        c[:, :, 0, 0], c[:, :, 0, 1] = 128, 96
        n_notes = rng.integers(1, 15, (4, t_len))
        pos = np.arange(16)[None, None, :]
        note = (pos >= 1) & (pos <= n_notes[..., None]) & s[..., None]
        c[..., 0] = np.where(note, rng.integers(0, 128, (4, t_len, 16)), c[..., 0])
        c[..., 1] = np.where(note, rng.integers(0, 96, (4, t_len, 16)), c[..., 1])
        eos_at = np.where(s, n_notes + 1, 1)
        tt, ss = np.meshgrid(np.arange(4), np.arange(t_len), indexing="ij")
        c[tt, ss, eos_at, 0], c[tt, ss, eos_at, 1] = 129, 97
        out.append((c, s))
    return out


def main():
    ref = ref_loader.load()
    cwd = os.getcwd()
    os.chdir(ref_loader.REFERENCE_DIR)
    try:
        utils = importlib.import_module("utils")
    finally:
        os.chdir(cwd)
    store = {}
    for n_bars in (2, 16):
        smp = samples(n_bars, seed=n_bars)
        with tempfile.TemporaryDirectory() as d:
            for k, (c, s) in enumerate(smp):
                np.savez(os.path.join(d, f"sample{k}"), c_tensor=c, s_tensor=s)
            ds = ref.data.PolyphemusDataset(d, n_bars=n_bars)
            names = [f.name for f in ds.files]
            for idx in range(len(ds)):
                k = int(names[idx].replace("sample", "").replace(".npz", ""))
                g = ds[idx]
                key = f"b{n_bars}.s{k}"
                store[key + ".c_disk"], store[key + ".s_disk"] = smp[k]
                store[key + ".s_tensor"] = g.s_tensor.numpy().astype(bool)
                ct = g.c_tensor.numpy()                                  # [N, 16, 230] one-hot
                assert ((ct[..., :131] == 1).sum(-1) == 1).all() and ((ct[..., 131:] == 1).sum(-1) == 1).all()
                store[key + ".tokens"] = np.stack((ct[..., :131].argmax(-1), ct[..., 131:].argmax(-1)), -1).astype(np.int16)
                store[key + ".edge_index"] = g.edge_index.numpy()
                store[key + ".num_nodes"] = np.int64(int(g.num_nodes))
    np.savez_compressed(os.path.join(HERE, "dataset_items.npz"), **store)
    # mtp_from_logits
    gen = torch.Generator().manual_seed(0)
    s_tensor = torch.rand(2, 2, 4, 32, generator=gen) < 0.08
    s_tensor[1, 0] = False
    n = int(s_tensor.sum())
    c_logits = torch.randn(n, 15, 230, generator=gen)
    mtp = utils.mtp_from_logits(c_logits, s_tensor)
    np.savez_compressed(os.path.join(HERE, "mtp_from_logits.npz"), s_tensor=s_tensor.numpy(), c_logits=c_logits.numpy(),
                        mtp_sum=mtp.numpy().sum(axis=-1), mtp_argmax=mtp.numpy().argmax(axis=-1).astype(np.int16))
    print("wrote dataset_items.npz, mtp_from_logits.npz:", {k: v.shape for k, v in store.items() if k.endswith("tokens")})


if __name__ == "__main__":
    main()
