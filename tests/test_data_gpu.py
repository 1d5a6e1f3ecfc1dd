"""Device-side dataset decode (data.py:218-271) and mtp_from_logits (utils.py:59-79): bit-exact against the golden
fixtures produced by the reference's own code and against the numpy oracle on fresh samples, through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import data_oracle as do
from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_bars", [2, 16])
def test_decode_samples_matches_reference_golden(cuda, n_bars):
    import polyphemus_b200 as pb

    ref = golden("dataset_items.npz")
    keys = [f"b{n_bars}.s{k}" for k in range(4)]
    c_disk = torch.from_numpy(np.stack([ref[k + ".c_disk"] for k in keys]))
    s_disk = torch.from_numpy(np.stack([ref[k + ".s_disk"] for k in keys]))
    g = pb.decode_samples(c_disk, s_disk, n_bars, device=cuda)
    want_s = np.stack([ref[k + ".s_tensor"] for k in keys])
    want_tok = np.concatenate([ref[k + ".tokens"] for k in keys])
    np.testing.assert_array_equal(g.s_tensor.cpu().numpy().astype(bool).reshape(want_s.shape), want_s)
    np.testing.assert_array_equal(g.c_tokens.cpu().numpy(), want_tok)
    offs = np.cumsum([0] + [int(ref[k + ".num_nodes"]) for k in keys])
    want_ei = np.concatenate([ref[k + ".edge_index"] + o for k, o in zip(keys, offs)], axis=1)
    np.testing.assert_array_equal(g.edge_index.cpu().numpy(), want_ei)
    assert g.num_nodes == offs[-1] and g.num_graphs == 4 and g.n_bars == n_bars
    # the one-hot view the reference attaches (data.py:233-259), on request
    g2 = pb.decode_samples(c_disk, s_disk, n_bars, device=cuda, onehot=True)
    np.testing.assert_array_equal(g2.c_tensor.cpu().numpy(), do.onehot(want_tok.astype(np.int64)))
    # a single sample, as dataset[i] hands it over
    g1 = pb.decode_samples(c_disk[1], s_disk[1], n_bars, device=cuda)
    np.testing.assert_array_equal(g1.c_tokens.cpu().numpy(), ref[keys[1] + ".tokens"])


def test_decode_samples_lmd16_batch_matches_oracle(cuda):
    """A training-size batch (LMD16, 64 samples, ~33k nodes) against the numpy oracle; pinned host input."""
    import polyphemus_b200 as pb
    from polyphemus_b200.train import synthetic_host_batch

    host = synthetic_host_batch(64, 16, 0.25, seed=3, pin=True, disk=True)
    g = pb.decode_samples(host.c_disk, host.s_disk, 16, device=cuda)
    s_all, tok_all = [], []
    for b in range(64):
        s, tok, _ = do.dataset_item(host.c_disk[b].numpy(), host.s_disk[b].numpy(), 16)
        s_all.append(s)
        tok_all.append(tok)
    np.testing.assert_array_equal(g.s_tensor.cpu().numpy().astype(bool).reshape(64, 16, 4, 32), np.stack(s_all))
    np.testing.assert_array_equal(g.c_tokens.cpu().numpy().astype(np.int64), np.concatenate(tok_all))
    np.testing.assert_array_equal(g.c_tokens.cpu().numpy(), host.tokens.numpy())       # == the compacted layout
    arrays = go.batch_graph(np.stack(s_all))
    np.testing.assert_array_equal(g.edge_index.cpu().numpy(), arrays.edge_index)
    with pytest.raises(ValueError):
        pb.decode_samples(host.c_disk, host.s_disk, 8, device=cuda)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mtp_from_logits_matches_reference_golden(cuda, dtype):
    import polyphemus_b200 as pb

    ref = golden("mtp_from_logits.npz")
    s = torch.from_numpy(ref["s_tensor"])
    logits = torch.from_numpy(ref["c_logits"]).to(dtype)
    mtp = pb.mtp_from_logits(logits.to(cuda), s.to(cuda))
    want = do.mtp_from_logits(logits.float().numpy(), ref["s_tensor"])
    assert mtp.dtype == dtype and tuple(mtp.shape) == want.shape
    np.testing.assert_array_equal(mtp.float().cpu().numpy(), want)
    if dtype == torch.float32:
        np.testing.assert_allclose(mtp.cpu().numpy().sum(-1), ref["mtp_sum"], rtol=0, atol=1e-5)
        np.testing.assert_array_equal(mtp.cpu().numpy().argmax(-1)[~ref["s_tensor"]], ref["mtp_argmax"][~ref["s_tensor"]])


def test_generation_end_to_end_to_pianoroll(cuda):
    """BASELINE config 3 in small: z ~ N(0, I) -> decoder with structure conditioning -> mtp_from_logits
    (generate.py:24-35, 226-237): shapes and the silence pattern of the reference."""
    import polyphemus_b200 as pb

    cfg = dict(dropout=0, batch_norm=True, gnn_n_layers=2, d=64, n_bars=2, resolution=8)
    torch.manual_seed(0)
    vae = pb.VAE(**cfg, device=cuda).to(cuda).eval()
    s_json = torch.from_numpy(golden("graph_structure_json.npz")["s_in"]).bool()[0]
    n = 8
    s_cond = s_json.unsqueeze(0).repeat(n, 1, 1, 1).to(cuda)
    with torch.no_grad():
        graph = vae.decoder._structure_from_binary(s_cond)
        z = torch.randn(n, cfg["d"], device=cuda)
        _, c_logits = vae.decoder(z, graph)
        mtp = pb.mtp_from_logits(c_logits, s_cond)
    assert tuple(mtp.shape) == (n, 2, 4, 32, 15, 230)
    assert torch.equal(mtp[s_cond], c_logits)
    silent = mtp[~s_cond]
    assert (silent.sum(-1) == 1).all() and (silent[:, 0].argmax(-1) == 129).all() and (silent[:, 1:].argmax(-1) == 130).all()
