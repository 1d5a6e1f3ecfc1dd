"""CPU tier for the product's host side: C-ABI surface, closed-form graph plan logic (host-compiled),
state-dict / seeded-init contract, loss restatement, data-parallel reducer over gloo (world size 2)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, golden
from oracle import graph_oracle as go
from oracle import model_oracle as mo


def test_abi_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "polyphemus_b200.h")).read()
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/polyphemus_b200.h but not exported"
    from polyphemus_b200 import _ffi

    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    lib.pb_version.restype = ctypes.c_int
    assert lib.pb_version() == 100


def test_kernels_are_native_blackwell(built_lib):
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")          # no legacy mma.sync path


def test_no_cpu_fallback():
    import polyphemus_b200 as pb

    with pytest.raises((pb.PolyphemusB200Error, RuntimeError)):
        s = torch.zeros(1, 2, 4, 32, dtype=torch.bool)
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; the no-device path cannot be exercised")
        pb.graphs_from_tensor(s)
    layer = pb.GCL(64, 64, 6, torch.nn.Linear(32, 64))
    with pytest.raises((pb.PolyphemusB200Error, RuntimeError)):
        layer(torch.randn(3, 64), torch.zeros(2, 1, dtype=torch.long), torch.zeros(1), torch.zeros(1, 32))


@pytest.fixture(scope="module")
def host_plan_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("host") / "libgraph_plan_host.so"
    src = os.path.join(ROOT, "tests", "host", "graph_plan_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), src], check=True)
    return ctypes.CDLL(str(out))


def _host_edges(lib, bar):
    bits = np.array([sum(int(bar[k, t]) << t for t in range(32)) for k in range(4)], dtype=np.uint32)
    out = np.zeros((1100, 4), dtype=np.int64)
    cnt = np.zeros(5, dtype=np.int32)
    n = lib.pbh_bar_edges(bits.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                          cnt.ctypes.data_as(ctypes.c_void_p))
    return out[:n], cnt


def test_graph_plan_logic_matches_oracle(host_plan_lib):
    """The integer code the CUDA builder runs per (bar, timestep) (csrc/graph_plan.h), compiled for the host."""
    rng = np.random.default_rng(0)
    bars = [rng.random((4, 32)) < p for p in (0.01, 0.03, 0.1, 0.25, 0.5, 0.8, 1.0) for _ in range(150)]
    ref = golden("graph_random.npz")
    bars += list(ref["edge.s_out"].reshape(-1, 4, 32))
    for bar in bars:
        bar = bar.copy()
        if not bar.any():
            bar[0, 0] = True
        want, n = go.bar_edges(bar)
        got, cnt = _host_edges(host_plan_lib, bar)
        assert cnt[0] == n and cnt[4] == want.shape[0]
        np.testing.assert_array_equal(got, want)


def test_state_dict_contract_and_seeded_init():
    import polyphemus_b200 as pb

    gold = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    torch.manual_seed(0)
    vae = pb.VAE(**gold["config"], device=torch.device("cpu"))
    sd = vae.state_dict()
    assert list(sd.keys()) == list(gold["keys"].keys())
    assert {k: list(v.shape) for k, v in sd.items()} == gold["keys"]
    assert sum(p.numel() for p in vae.parameters()) == gold["n_params"] == 42210909
    lmd16 = pb.VAE(**{**gold["config"], "n_bars": 16}, device=torch.device("cpu"))
    assert sum(p.numel() for p in lmd16.parameters()) == 56905309
    # the edge network is one module shared by all layers of a GCN (model.py:175)
    enc = vae.encoder.c_encoder.graph_encoder
    assert all(layer.nn is enc.layers[0].nn for layer in enc.layers)


def test_seeded_init_equals_golden_parameters():
    import polyphemus_b200 as pb

    ref = golden("vae_step.npz")
    cfg = json.loads(str(ref["config"]))
    torch.manual_seed(3)                                   # make_golden.py seeds the reference VAE with 3
    vae = pb.VAE(**cfg, device=torch.device("cpu"))
    sd = vae.state_dict()
    for k in ("encoder.c_encoder.graph_encoder.layers.1.weight", "decoder.c_decoder.graph_decoder.layers.0.root",
              "encoder.c_encoder.graph_encoder.layers.0.nn.weight", "decoder.c_decoder.chord_decoder.weight"):
        assert torch.equal(sd[k], torch.from_numpy(ref["sd." + k])), k


def test_loss_restatement_matches_oracle():
    from polyphemus_b200.train import onehot_content, synthetic_tokens, vae_losses

    g = torch.Generator().manual_seed(0)
    n, b = 37, 3
    tokens = synthetic_tokens(n, g)
    c_tensor = onehot_content(tokens)
    torch.testing.assert_close(c_tensor, mo.onehot_content(tokens.long()))
    c_logits = torch.randn(n, 15, 230, generator=g)
    s_tensor = (torch.rand(b * 2, 4, 32, generator=g) < 0.25).float()
    s_logits = torch.randn(b, 2, 4, 32, generator=g)
    mu, log_var = torch.randn(b, 8, generator=g), torch.randn(b, 8, generator=g)
    want, wp = mo.losses(s_tensor, s_logits, c_tensor, c_logits, mu, log_var, beta=0.3)
    for toks in (None, tokens):
        got, gp = vae_losses(s_tensor, s_logits, c_tensor, c_logits, mu, log_var, beta=0.3, c_tokens=toks)
        torch.testing.assert_close(got, want)
        for k in wp:
            torch.testing.assert_close(gp[k], wp[k])


def test_dropout_rng_known_answer():
    """The dropout mask hashes a counter with the SplitMix64 output function (csrc/common.cuh). Known answers:
    the first outputs of SplitMix64 seeded with 0 and with 1234567 (Vigna's reference implementation)."""
    M = (1 << 64) - 1

    def mix(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    def splitmix_stream(seed, n):
        out, x = [], seed
        for _ in range(n):
            x = (x + 0x9E3779B97F4A7C15) & M
            out.append(mix(x))
        return out

    assert splitmix_stream(0, 3) == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    assert splitmix_stream(1234567, 2) == [6457827717110365317, 3203168211198807973]

    # our counter layout: bits(seed, eid, chunk) = mix(seed + GOLDEN * (1 + (eid << 32 | chunk))); with eid = 0 the
    # chunk-th value of the stream seeded with `seed` — i.e. exactly SplitMix64.
    def bits(seed, eid, chunk):
        return mix((seed + 0x9E3779B97F4A7C15 * (((eid << 32) | chunk) + 1)) & M)

    assert [bits(0, 0, c) for c in range(3)] == splitmix_stream(0, 3)
    keep = [((bits(42, 7, 3) >> (16 * i)) & 0xFFFF) >= 6554 for i in range(4)]      # p = 0.1 -> thresh 6554
    assert len(keep) == 4


DP_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from polyphemus_b200.train import GradAllReducer
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
unused = model[3]                      # never used in forward: grad stays None (cf. s_decoder, training.py:307)
red = GradAllReducer(model.parameters(), bucket_mb=0.0001)
data = torch.randn(world * 6, 8, generator=torch.Generator().manual_seed(1))
target = torch.randn(world * 6, 4, generator=torch.Generator().manual_seed(2))
for step in range(3):
    red.zero_grad()
    shard = slice(rank * 6, (rank + 1) * 6)
    loss = ((model[2](model[1](model[0](data[shard]))) - target[shard]) ** 2).mean()
    loss.backward()
    red.finish()
    # the unused layer is registered last = first buckets: before the first finish() has agreed on the gradient-less
    # parameters nothing can go out during backward, afterwards every bucket does
    assert red.launched_in_backward == (0 if step == 0 else len(red.buckets)), (step, red.launched_in_backward)
assert len(red._absent) == 2
# single-process reference: average of per-shard gradients
ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
ref.load_state_dict(model.state_dict())
grads = None
for r in range(world):
    ref.zero_grad()
    shard = slice(r * 6, (r + 1) * 6)
    ((ref[2](ref[1](ref[0](data[shard]))) - target[shard]) ** 2).mean().backward()
    g = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in ref.parameters()]
    grads = g if grads is None else [a + b for a, b in zip(grads, g)]
for p, g in zip(model.parameters(), grads):
    assert torch.allclose(p.grad, g / world, atol=1e-6), (rank, p.shape)
assert float(unused.weight.grad.abs().max()) == 0.0
assert len(red.buckets) > 1
print("rank", rank, "ok")
"""


def test_grad_all_reducer_world_size_2_gloo(tmp_path):
    script = tmp_path / "dp_check.py"
    script.write_text(DP_SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_batched_bn_tables_match_per_table_path():
    """ContentEncoder._bn_tables_batched (one batched computation for the four BatchNorm(Linear(one_hot)) tables of a
    step) == four `_bn_table` calls in the reference's order: tables, gradients and running statistics, including a
    token set that is empty (no drum nodes)."""
    import copy

    import polyphemus_b200 as pb
    from polyphemus_b200.vae import N_DUR_TOKENS, N_PITCH_TOKENS

    torch.manual_seed(0)
    cfg = dict(dropout=0, batch_norm=True, gnn_n_layers=1, d=64, n_bars=2, resolution=8)
    enc_a = pb.ContentEncoder(**cfg).train()
    with torch.no_grad():
        for p in enc_a.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    enc_b = copy.deepcopy(enc_a)
    gen = torch.Generator().manual_seed(1)
    for empty_drums in (False, True):
        cnt_p = torch.randint(0, 50, (2, N_PITCH_TOKENS), generator=gen)
        cnt_d = torch.randint(0, 50, (2, N_DUR_TOKENS), generator=gen)
        if empty_drums:
            cnt_p[1], cnt_d[1] = 0, 0
        p_tabs, d_tabs = enc_a._bn_tables_batched(cnt_p, cnt_d)
        ref_p1 = enc_b._bn_table(enc_b.drums_pitch_emb, enc_b.bn_drums, None, True, cnt_p[1])
        ref_d1 = enc_b._bn_table(enc_b.dur_emb, enc_b.bn_dur, None, True, cnt_d[1])
        ref_p0 = enc_b._bn_table(enc_b.non_drums_pitch_emb, enc_b.bn_non_drums, None, True, cnt_p[0])
        ref_d0 = enc_b._bn_table(enc_b.dur_emb, enc_b.bn_dur, None, True, cnt_d[0])
        for got, want in ((p_tabs[0], ref_p0), (p_tabs[1], ref_p1), (d_tabs[0], ref_d0), (d_tabs[1], ref_d1)):
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
        w = torch.randn(2, N_PITCH_TOKENS, 32, generator=gen), torch.randn(2, N_DUR_TOKENS, 32, generator=gen)
        ((p_tabs * w[0]).sum() + (d_tabs * w[1]).sum()).backward()
        ((torch.stack((ref_p0, ref_p1)) * w[0]).sum() + (torch.stack((ref_d0, ref_d1)) * w[1]).sum()).backward()
    for (name, pa), (_, pb_) in zip(enc_a.named_parameters(), enc_b.named_parameters()):
        if pa.grad is None:
            assert pb_.grad is None, name
            continue
        # (a bias in front of a BatchNorm has a mathematically zero gradient whenever its token set is not empty: both
        # sides then hold round-off only, hence the absolute tolerance at the scale of the weight gradient)
        ref_w = dict(enc_b.named_parameters()).get(name.replace(".bias", ".weight"))
        scale = float(pb_.grad.abs().max()) if ref_w is None or ref_w.grad is None else float(ref_w.grad.abs().max())
        torch.testing.assert_close(pa.grad, pb_.grad, rtol=1e-4, atol=1e-5 * max(scale, 1e-30), msg=name)
    sa, sb = enc_a.state_dict(), enc_b.state_dict()
    for k in sa:
        if "running" in k or "num_batches" in k:
            torch.testing.assert_close(sa[k].float(), sb[k].float(), rtol=1e-5, atol=1e-6, msg=k)
