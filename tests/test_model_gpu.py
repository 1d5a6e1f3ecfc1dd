"""Module-level GPU parity: GCL / GCN / whole VAE training step against the golden fixtures produced by the
reference's own code (tests/golden/make_golden.py) and against the oracle restatement on fresh seeded inputs.

fp32 mode tolerance: rtol 1e-4 / atol 1e-5 (BASELINE.json north_star), GCL-internal dropout 0 on both sides
(SURVEY.md §7); the dropout path is checked exactly through the exported Philox keep-mask.
"""
import json

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import graph_oracle as go
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-4, atol=1e-5)


def _t(a, cuda):
    return torch.from_numpy(np.asarray(a)).to(cuda)


def _gcl_from_golden(ref, cuda, precision):
    import polyphemus_b200 as pb

    d = ref["x"].shape[1]
    edge_nn = torch.nn.Linear(32, d)
    layer = pb.GCL(d, d, 6, edge_nn, dropout=0.0, precision=precision).to(cuda)
    with torch.no_grad():
        layer.weight.copy_(_t(ref["weight"], cuda))
        layer.root.copy_(_t(ref["root"], cuda))
        layer.bias.copy_(_t(ref["bias"], cuda))
        layer.nn.weight.copy_(_t(ref["nn_weight"], cuda))
        layer.nn.bias.copy_(_t(ref["nn_bias"], cuda))
    return layer


def _edge_inputs(ref, cuda):
    et = torch.from_numpy(ref["edge_type"])
    ed = torch.from_numpy(ref["edge_dist"])
    edge_attr = torch.nn.functional.one_hot(ed, 32).float()
    return _t(ref["edge_index"], cuda), et.float().to(cuda), edge_attr.to(cuda)


def test_gcl_matches_reference_golden_fp32(cuda):
    """PyG-style call conv(x, edge_index, edge_type, edge_attr) (model.py:55-57,200), forward + all gradients."""
    ref = golden("gcl_layer.npz")
    layer = _gcl_from_golden(ref, cuda, "fp32")
    layer.train()
    ei, et, ea = _edge_inputs(ref, cuda)
    x = _t(ref["x"], cuda).requires_grad_(True)
    y = layer(x, ei, et, ea)
    torch.testing.assert_close(y.detach().cpu(), torch.from_numpy(ref["y"]), **TOL)
    y.backward(_t(ref["gy"], cuda))
    for got, name in ((x.grad, "gx"), (layer.weight.grad, "g_weight"), (layer.root.grad, "g_root"),
                      (layer.bias.grad, "g_bias"), (layer.nn.weight.grad, "g_nn_weight"), (layer.nn.bias.grad, "g_nn_bias")):
        want = torch.from_numpy(ref[name])
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5 * max(1.0, float(want.abs().max())), msg=name)


def test_gcl_bf16_within_bf16_budget(cuda):
    ref = golden("gcl_layer.npz")
    layer = _gcl_from_golden(ref, cuda, "bf16")
    ei, et, ea = _edge_inputs(ref, cuda)
    x = _t(ref["x"], cuda).requires_grad_(True)
    y = layer(x, ei, et, ea)
    want = torch.from_numpy(ref["y"])
    err = (y.detach().cpu() - want).abs().max() / want.abs().max()
    assert err < 2e-2, f"bf16 forward relative-to-scale error {err}"
    y.backward(_t(ref["gy"], cuda))
    for got, name in ((x.grad, "gx"), (layer.weight.grad, "g_weight"), (layer.root.grad, "g_root")):
        want = torch.from_numpy(ref[name])
        err = (got.cpu() - want).abs().max() / want.abs().max()
        assert err < 3e-2, f"bf16 {name} relative-to-scale error {err}"


def _gcn_from_golden(ref, cuda, precision="fp32"):
    import polyphemus_b200 as pb

    d = ref["x"].shape[1]
    gcn = pb.GCN(input_dim=d, hidden_dim=d, n_layers=2, num_relations=6, batch_norm=True, dropout=0, precision=precision)
    sd = {k[3:]: torch.from_numpy(ref[k]) for k in ref.files if k.startswith("sd.")}
    gcn.load_state_dict(sd)
    for layer in gcn.layers:
        layer.dropout = 0.0
    return gcn.to(cuda)


class _Data:
    pass


def test_gcn_stack_matches_reference_golden(cuda):
    """GCN(data) reading .x/.edge_index/.edge_attrs (model.py:192): forward, running stats, gradients, eval."""
    ref = golden("gcn_stack.npz")
    gcn = _gcn_from_golden(ref, cuda)
    gcn.train()
    data = _Data()
    data.x = _t(ref["x"], cuda).requires_grad_(True)
    data.edge_index = _t(ref["edge_index"], cuda)
    arrays_attrs = np.zeros((ref["edge_type"].shape[0], 33), dtype=np.float32)
    arrays_attrs[:, 0] = ref["edge_type"]
    arrays_attrs[np.arange(arrays_attrs.shape[0]), ref["edge_dist"] + 1] = 1
    data.edge_attrs = _t(arrays_attrs, cuda)
    y = gcn(data)
    torch.testing.assert_close(y.detach().cpu(), torch.from_numpy(ref["y"]), **TOL)
    y.backward(_t(ref["gy"], cuda))
    torch.testing.assert_close(data.x.grad.cpu(), torch.from_numpy(ref["gx"]), rtol=1e-4, atol=1e-5)
    for name, p in gcn.named_parameters():
        want = torch.from_numpy(ref["grad." + name])
        atol = 1e-5 * max(1.0, float(want.abs().max()))
        if name.endswith(".bias") and "norm_layers" not in name:
            # a bias in front of BatchNorm has a mathematically zero gradient: both sides hold only the round-off
            # of a cancelling column sum over |g_out| ~ 10, i.e. a few 1e-5
            atol = 3e-4
        torch.testing.assert_close(p.grad.cpu(), want, rtol=1e-4, atol=atol, msg=name)
    after = gcn.state_dict()
    for k in ref.files:
        if k.startswith("sd_after."):
            torch.testing.assert_close(after[k[9:]].cpu().float(), torch.from_numpy(ref[k]).float(), rtol=1e-4, atol=1e-5, msg=k)
    gcn.eval()
    with torch.no_grad():
        data.x = data.x.detach()
        y_eval = gcn(data)
    torch.testing.assert_close(y_eval.cpu(), torch.from_numpy(golden("gcn_stack_eval.npz")["y_eval"]), **TOL)


def test_gcl_dropout_path_is_exact_given_the_mask(cuda):
    """Training-mode GCL with p=0.1: the oracle driven by the kernel's own Philox keep-mask must agree (fwd+bwd)."""
    from polyphemus_b200 import ops
    import polyphemus_b200 as pb

    ref = golden("gcl_layer.npz")
    layer = _gcl_from_golden(ref, cuda, "fp32")
    layer.dropout = 0.1
    layer.train()
    ei, et, ea = _edge_inputs(ref, cuda)
    plan = layer.plan_from(_t(ref["x"], cuda), ei, et, ea)
    seed = 987654321
    x = _t(ref["x"], cuda).requires_grad_(True)
    y = ops.rgc_layer(x, layer.weight, layer.root, layer.bias, layer.nn.weight, layer.nn.bias, plan, batch_norm=False,
                      training=True, p_drop=0.1, precision="fp32", seed=seed)
    y.backward(_t(ref["gy"], cuda))
    keep = ops.dropout_keep_mask(ei.shape[1], x.shape[1], 0.1, seed, cuda).cpu()
    assert 0.88 < keep.float().mean() < 0.92
    xo = torch.from_numpy(ref["x"]).requires_grad_(True)
    params = [torch.from_numpy(ref[k]).clone().requires_grad_(True) for k in ("weight", "root", "bias", "nn_weight", "nn_bias")]
    yo = mo.gcl_forward(xo, torch.from_numpy(ref["edge_index"]), torch.from_numpy(ref["edge_type"]),
                        torch.from_numpy(ref["edge_dist"]), *params, keep_mask=keep, p_drop=0.1)
    yo.backward(torch.from_numpy(ref["gy"]))
    torch.testing.assert_close(y.detach().cpu(), yo.detach(), **TOL)
    torch.testing.assert_close(x.grad.cpu(), xo.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.weight.grad.cpu(), params[0].grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(layer.nn.weight.grad.cpu(), params[3].grad, rtol=1e-4, atol=1e-4)
    # a different seed gives a different mask; eval mode ignores dropout entirely
    assert not torch.equal(keep, ops.dropout_keep_mask(ei.shape[1], x.shape[1], 0.1, seed + 1, cuda).cpu())
    layer.eval()
    with torch.no_grad():
        torch.testing.assert_close(layer(x.detach(), ei, et, ea).cpu(), torch.from_numpy(ref["y"]), **TOL)


def _vae_from_golden(ref, cuda, precision):
    import polyphemus_b200 as pb

    cfg = json.loads(str(ref["config"]))
    pb.set_precision(precision)
    vae = pb.VAE(**cfg, device=cuda)
    sd = {k[3:]: torch.from_numpy(ref[k]) for k in ref.files if k.startswith("sd.")}
    vae.load_state_dict(sd)
    for m in vae.modules():
        if isinstance(m, pb.GCL):
            m.dropout = 0.0
    return vae.to(cuda), cfg


def _vae_batch(ref, cuda):
    from polyphemus_b200.train import HostBatch, device_batch

    host = HostBatch(torch.from_numpy(ref["s_in"].copy()), torch.from_numpy(ref["tokens"].copy()))
    return device_batch(host, cuda, onehot=True)


@pytest.mark.parametrize("content", ["token_ids", "onehot", "token_ids_lazy_logits", "token_ids_lazy_logits_onematrix",
                                     "token_ids_lazy_unfolded"])
def test_vae_training_step_matches_reference_golden(cuda, content):
    """Whole drop-in surface: VAE(graph) -> ((s_logits, c_logits), mu, log_var), reference loss, all gradients,
    BatchNorm running statistics; inputs go through the device graph builder (one empty bar included).
    `content`: note tokens as ids (dataset layout, token-table embedding path) or as the reference's one-hot
    float c_tensor (data.py:234-259, generic Linear + BatchNorm path)."""
    from polyphemus_b200.train import vae_losses
    import polyphemus_b200 as pb

    ref = golden("vae_step.npz")
    try:
        vae, cfg = _vae_from_golden(ref, cuda, "fp32")
        vae.train()
        graph = _vae_batch(ref, cuda)
        tokens = graph.c_tokens
        if content == "onehot":
            graph.c_tokens = None
        if content.startswith("token_ids_lazy"):     # what train.TrainStep does: loss straight from the head outputs
            vae.decoder.c_decoder.materialize_logits = False
        if content == "token_ids_lazy_logits_onematrix":   # PB200_SPLIT_HEADS=0: one [drum | non-drum | duration] matrix for all rows
            pb.ops._split_heads = False
        if content == "token_ids_lazy_unfolded":
            # an active dropout between chord_decoder and the heads forbids composing them; p = 1e-12 keeps every
            # element and scales by exactly 1.0f, so the separate-heads path must reproduce the same golden numbers
            vae.decoder.c_decoder.dropout_layer.p = 1e-12
        (s_logits, c_logits), mu, log_var = vae(graph, noise=_t(ref["noise"], cuda))
        c_parts = c_logits
        if content.startswith("token_ids_lazy"):
            assert isinstance(c_logits, pb.vae.LogitParts)
            assert (c_parts.split is not None) == (content == "token_ids_lazy_logits")
            assert (c_parts.combined is not None) == (content == "token_ids_lazy_logits_onematrix")
            c_logits = c_parts.dense()
        # mu / log_var sit behind a BatchNorm over a batch of 4 sequences: reference self-noise level
        torch.testing.assert_close(mu.detach().cpu(), torch.from_numpy(ref["mu"]), rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(log_var.detach().cpu(), torch.from_numpy(ref["log_var"]), rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(s_logits.detach().cpu(), torch.from_numpy(ref["s_logits"]), **TOL)
        torch.testing.assert_close(c_logits.detach().cpu(), torch.from_numpy(ref["c_logits"]), rtol=1e-4, atol=2e-5)
        for use_tokens in (False, True):
            loss, parts = vae_losses(graph.s_tensor, s_logits, graph.c_tensor, c_parts, mu, log_var, beta=0.0,
                                     c_tokens=tokens if use_tokens else None)
            assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
        got_parts = torch.stack([parts[k] for k in ("pitch", "dur", "structure", "kld")]).detach().cpu()
        torch.testing.assert_close(got_parts, torch.from_numpy(ref["loss_parts"]), rtol=1e-4, atol=1e-5)
        loss.backward()
        n_none = 0
        for name, p in vae.named_parameters():
            want = torch.from_numpy(ref["grad." + name])
            if want.numel() == 0:                       # reference grad is None (s_decoder, training.py:307)
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                n_none += 1
                continue
            torch.testing.assert_close(p.grad.cpu(), want, rtol=1e-4, atol=1e-5 * max(1.0, float(want.abs().max())), msg=name)
        assert n_none == 12
        after = vae.state_dict()
        for k in ref.files:
            if k.startswith("sd_after."):
                torch.testing.assert_close(after[k[9:]].cpu().float(), torch.from_numpy(ref[k]).float(), rtol=1e-4,
                                           atol=1e-5, msg=k)
    finally:
        pb.set_precision("fp32")
        pb.ops._split_heads = True


def test_vae_bf16_mode_tracks_fp32(cuda):
    import polyphemus_b200 as pb
    from polyphemus_b200.train import vae_losses

    ref = golden("vae_step.npz")
    try:
        vae, _ = _vae_from_golden(ref, cuda, "bf16")
        vae.train()
        graph = _vae_batch(ref, cuda)
        (s_logits, c_logits), mu, log_var = vae(graph, noise=_t(ref["noise"], cuda))
        loss, _ = vae_losses(graph.s_tensor, s_logits, None, c_logits, mu, log_var, c_tokens=graph.c_tokens)
        assert abs(float(loss) - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"]))
        loss.backward()
        for key, mod in (("decoder.c_decoder.graph_decoder.layers.1.weight", vae.decoder.c_decoder.graph_decoder.layers[1].weight),
                         ("decoder.c_decoder.graph_decoder.layers.0.root", vae.decoder.c_decoder.graph_decoder.layers[0].root),
                         ("decoder.c_decoder.chord_decoder.weight", vae.decoder.c_decoder.chord_decoder.weight)):
            want = torch.from_numpy(ref["grad." + key]).flatten().double()
            got = mod.grad.cpu().flatten().double()
            cos = float(torch.dot(got, want) / (got.norm() * want.norm()))
            assert cos > 0.99, f"bf16 gradient direction of {key}: cos={cos}"
            assert 0.9 < float(got.norm() / want.norm()) < 1.1, key
    finally:
        pb.set_precision("fp32")


def test_vae_against_oracle_on_fresh_inputs(cuda):
    """Independent of the fixtures: random init here, oracle restatement on CPU, LMD2-like shape."""
    import polyphemus_b200 as pb
    from polyphemus_b200.train import HostBatch, device_batch, synthetic_tokens, vae_losses

    cfg = dict(dropout=0, batch_norm=True, gnn_n_layers=3, d=128, n_bars=2, resolution=8)
    torch.manual_seed(5)
    vae = pb.VAE(**cfg, device=cuda)
    for m in vae.modules():
        if isinstance(m, pb.GCL):
            m.dropout = 0.0
    with torch.no_grad():
        for p in vae.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd_cpu = {k: v.detach().clone() for k, v in vae.state_dict().items()}
    vae = vae.to(cuda).train()
    s_np = go.synthetic_structure(6, 2, 0.25, seed=21)
    arrays = go.batch_graph(s_np)
    tokens = synthetic_tokens(arrays.num_nodes, torch.Generator().manual_seed(3))
    noise = torch.randn(6, cfg["d"], generator=torch.Generator().manual_seed(4))
    graph = device_batch(HostBatch(torch.from_numpy(arrays.s_tensor.copy()), tokens), cuda)
    (s_logits, c_logits), mu, log_var = vae(graph, noise=noise.to(cuda))
    loss, _ = vae_losses(graph.s_tensor, s_logits, None, c_logits, mu, log_var, c_tokens=graph.c_tokens)
    loss.backward()
    sd = mo.leaf_state(sd_cpu)
    gb = mo.make_batch(arrays, tokens.long())
    ctx = mo.Ctx(training=True)
    (s2, c2), mu2, lv2 = mo.vae(sd, gb, cfg["n_bars"], cfg["d"], ctx, eps_noise=noise)
    loss2, _ = mo.losses(gb.s_tensor, s2, gb.c_tensor, c2, mu2, lv2)
    loss2.backward()
    torch.testing.assert_close(c_logits.detach().cpu(), c2.detach(), rtol=1e-4, atol=2e-5)
    assert abs(float(loss) - float(loss2)) <= 1e-4 * abs(float(loss2))
    for name, p in vae.named_parameters():
        want = sd[name].grad
        if want is None:
            continue
        torch.testing.assert_close(p.grad.cpu(), want, rtol=1e-4, atol=1e-5 * max(1.0, float(want.abs().max())), msg=name)


def test_generation_path_decoder_only(cuda):
    """generate.py:24-35 + 226-237: z ~ N(0,I) -> decoder(z, s) with structure conditioning, and unconditioned
    (decoder builds the graph from its own thresholded logits, model.py:646-650), eval mode."""
    import os
    import polyphemus_b200 as pb

    cfg = dict(dropout=0, batch_norm=True, gnn_n_layers=2, d=64, n_bars=2, resolution=8)
    torch.manual_seed(0)
    vae = pb.VAE(**cfg, device=cuda).to(cuda).eval()
    s_json = torch.from_numpy(golden("graph_structure_json.npz")["s_in"]).bool()[0]
    n = 16
    s_cond = s_json.unsqueeze(0).repeat(n, 1, 1, 1).to(cuda)
    with torch.no_grad():
        s = vae.decoder._structure_from_binary(s_cond)
        assert s.num_nodes == n * 30
        z = torch.randn(n, cfg["d"], device=cuda)
        s_logits, c_logits = vae.decoder(z, s)
        assert c_logits.shape == (n * 30, 15, 230) and s_logits.shape == (n, 2, 4, 32)
        assert torch.isfinite(c_logits).all()
        s_logits2, c_logits2 = vae.decoder(z)                         # unconditioned
        binary = vae.decoder._binary_from_logits(s_logits2)
        assert binary.flatten(-2).any(-1).all()                       # no empty bars after the fake activation
        assert c_logits2.shape[0] == int(binary.sum())


def test_layer_is_bit_reproducible(cuda):
    ref = golden("gcn_stack.npz")
    outs = []
    for _ in range(2):
        gcn = _gcn_from_golden(ref, cuda).train()
        data = _Data()
        data.x = _t(ref["x"], cuda).requires_grad_(True)
        data.edge_index = _t(ref["edge_index"], cuda)
        attrs = np.zeros((ref["edge_type"].shape[0], 33), dtype=np.float32)
        attrs[:, 0] = ref["edge_type"]
        attrs[np.arange(attrs.shape[0]), ref["edge_dist"] + 1] = 1
        data.edge_attrs = _t(attrs, cuda)
        y = gcn(data)
        y.backward(_t(ref["gy"], cuda))
        outs.append((y.detach().clone(), data.x.grad.clone(), gcn.layers[0].weight.grad.clone(),
                     gcn.layers[0].nn.weight.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def _structured_case(cuda, d=256, precision="fp32", p_gcl=0.0, seed=0):
    """A GCN stack on a builder graph with every special case of the structured layout: an empty bar (fake node),
    single-node bars on tracks 0 and 2 (group 0 through the fake self-edge), a full bar, empty groups elsewhere."""
    import polyphemus_b200 as pb

    s_np = go.synthetic_structure(5, 3, 0.3, seed=seed)
    s_np[0, 1] = False                       # empty bar
    s_np[1, 0] = False
    s_np[1, 0, 2, 9] = True                  # one node, track 2
    s_np[2, 2] = False
    s_np[2, 2, 0, 0] = True                  # one node, track 0
    s_np[3, 1] = True                        # full bar
    arrays = go.batch_graph(s_np)
    graph = pb.graphs_from_tensor(torch.from_numpy(s_np).to(cuda))
    torch.manual_seed(seed)
    gcn = pb.GCN(input_dim=d, hidden_dim=d, n_layers=2, num_relations=6, batch_norm=True, dropout=0, precision=precision)
    for layer in gcn.layers:
        layer.dropout = p_gcl
    with torch.no_grad():
        for prm in gcn.parameters():
            if prm.dim() == 1:
                prm.add_(0.1 * torch.randn_like(prm))
    sd = {k: v.detach().clone() for k, v in gcn.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(arrays.num_nodes, d, generator=gen)
    gy = torch.randn(arrays.num_nodes, d, generator=gen)
    return gcn.to(cuda).train(), sd, graph, arrays, x, gy


def _run_gcn(gcn, graph, x, gy, cuda):
    for prm in gcn.parameters():
        prm.grad = None
    graph.x = x.to(cuda).requires_grad_(True)
    y = gcn(graph)
    y.backward(gy.to(cuda))
    grads = {k: p.grad.detach().cpu().clone() for k, p in gcn.named_parameters()}
    stats = {k: v.detach().cpu().clone() for k, v in gcn.state_dict().items() if "running" in k}
    return y.detach().cpu(), graph.x.grad.cpu(), grads, stats


def test_structured_layout_matches_generic_and_oracle(cuda):
    """Track-relation-sorted 4d-wide operand == generic 7d-wide path == oracle, fwd + all gradients."""
    from polyphemus_b200 import ops

    gcn, sd, graph, arrays, x, gy = _structured_case(cuda)
    st = graph.structured
    assert st is not None and sum(st.counts) == arrays.num_nodes and st.n_padded % 128 == 0
    # group of a node: its track, except the lone node of a one-node bar (fake self-edge of type 0)
    grp = graph.node_group.cpu().numpy()
    bar_id = (graph.bars + 3 * graph.batch).cpu().numpy()
    lone = np.bincount(bar_id)[bar_id] == 1
    np.testing.assert_array_equal(grp, np.where(lone, 0, arrays.node_features.argmax(1)))
    try:
        ops.set_structured(True)
        y_s, gx_s, g_s, st_s = _run_gcn(gcn, graph, x, gy, cuda)
        gcn.load_state_dict(sd)
        ops.set_structured(False)
        y_g, gx_g, g_g, st_g = _run_gcn(gcn, graph, x, gy, cuda)
    finally:
        ops.set_structured(True)
    torch.testing.assert_close(y_s, y_g, **TOL)
    torch.testing.assert_close(gx_s, gx_g, rtol=1e-4, atol=1e-5)
    for k in g_g:
        atol = 3e-4 if (k.endswith(".bias") and "norm_layers" not in k) else 1e-5 * max(1.0, float(g_g[k].abs().max()))
        torch.testing.assert_close(g_s[k], g_g[k], rtol=1e-4, atol=atol, msg=k)
    for k in st_g:
        torch.testing.assert_close(st_s[k], st_g[k], rtol=1e-5, atol=1e-6, msg=k)
    # oracle
    osd = mo.leaf_state({"g." + k: v for k, v in sd.items()})
    xo = x.clone().requires_grad_(True)
    yo = mo.gcn_forward(osd, "g", xo, torch.from_numpy(arrays.edge_index), torch.from_numpy(arrays.edge_type),
                        torch.from_numpy(arrays.edge_dist), mo.Ctx(training=True))
    yo.backward(gy)
    torch.testing.assert_close(y_s, yo.detach(), **TOL)
    torch.testing.assert_close(gx_s, xo.grad, rtol=1e-4, atol=1e-5)
    for k in ("layers.0.weight", "layers.1.root", "layers.0.nn.weight", "norm_layers.1.module.weight"):
        want = osd["g." + k].grad
        torch.testing.assert_close(g_s[k], want, rtol=1e-4, atol=1e-5 * max(1.0, float(want.abs().max())), msg=k)


def test_structured_layout_bf16_and_dropout(cuda):
    """bf16 mode + GCL dropout 0.1 through the structured layout: same dropout decisions as the generic layout
    (they are keyed by the original edge id), results equal to bf16 round-off."""
    from polyphemus_b200 import ops
    import polyphemus_b200.ops as ops_mod
    import itertools

    gcn, sd, graph, arrays, x, gy = _structured_case(cuda, precision="bf16", p_gcl=0.1, seed=3)
    outs = []
    try:
        for flag in (True, False):
            ops.set_structured(flag)
            gcn.load_state_dict(sd)
            torch.manual_seed(11)
            ops_mod._seed_counter = itertools.count()          # same per-layer dropout seeds in both runs
            outs.append(_run_gcn(gcn, graph, x, gy, cuda))
    finally:
        ops.set_structured(True)
    (y_s, gx_s, g_s, _), (y_g, gx_g, g_g, _) = outs
    assert (y_s - y_g).abs().max() / y_g.abs().max() < 2e-2
    # the structured stack stores its activations in bf16: a pre-activation within bf16 round-off of the ReLU kink may
    # take the other branch, which moves single gradient entries by O(1) — bound the error in norm, and its maximum
    # loosely
    assert float((gx_s - gx_g).norm() / gx_g.norm()) < 3e-2
    assert (gx_s - gx_g).abs().max() / gx_g.abs().max() < 0.25
    for k in ("layers.0.weight", "layers.1.root", "layers.0.nn.weight"):
        a, b = g_s[k].flatten().double(), g_g[k].flatten().double()
        assert float(torch.dot(a, b) / (a.norm() * b.norm())) > 0.999, k
