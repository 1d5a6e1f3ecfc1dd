"""GPU parity of the device graph builder and the CSR plan against the oracle and the golden fixtures
(bit-exact: integer work). Calls go through the C ABI (polyphemus_b200._ffi)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu


def _build(s_np, cuda, **kw):
    import polyphemus_b200 as pb

    s = torch.from_numpy(np.array(s_np, dtype=bool, copy=True)).to(cuda)
    g = pb.graphs_from_tensor(s, **kw) if s.dim() == 4 else pb.graph_from_tensor(s)
    return g, s


def _assert_same(g, s_dev, ref, prefix=""):
    get = (lambda k: ref[prefix + k]) if not isinstance(ref, go.GraphArrays) else (lambda k: getattr(ref, k))
    assert g.num_nodes == int(get("num_nodes"))
    np.testing.assert_array_equal(g.edge_index.cpu().numpy(), get("edge_index"))
    np.testing.assert_array_equal(g.edge_type.cpu().numpy().astype(np.int64), get("edge_type"))
    np.testing.assert_array_equal(g.edge_dist.cpu().numpy().astype(np.int64), get("edge_dist"))
    np.testing.assert_array_equal(g.node_features.cpu().numpy(), get("node_features"))
    np.testing.assert_array_equal(g.is_drum.cpu().numpy(), get("is_drum"))
    np.testing.assert_array_equal(g.bars.cpu().numpy(), get("bars"))
    np.testing.assert_array_equal(g.batch.cpu().numpy(), get("batch"))
    s_expected = get("s_tensor") if isinstance(ref, go.GraphArrays) else get("s_out")
    np.testing.assert_array_equal(s_dev.cpu().numpy().reshape(s_expected.shape), s_expected)


def test_structure_json_golden(cuda):
    ref = golden("graph_structure_json.npz")
    g, s = _build(ref["s_in"], cuda)
    _assert_same(g, s, ref)
    # the known answers quoted in SURVEY.md §8c (bar 0 of structure.json)
    ei = g.edge_index.cpu().numpy()
    et, ed = g.edge_type.cpu().numpy(), g.edge_dist.cpu().numpy()
    first = np.stack([ei[0, :4], ei[1, :4], et[:4], ed[:4]], 1).tolist()
    assert first == [[0, 1, 0, 16], [1, 2, 0, 8], [1, 0, 0, 16], [2, 1, 0, 8]]


@pytest.mark.parametrize("case", ["bern0", "bern1", "bern2", "bern3", "bern4", "edge", "lmd16"])
def test_random_golden(cuda, case):
    ref = golden("graph_random.npz")
    g, s = _build(ref[f"{case}.s_in"], cuda)
    _assert_same(g, s, ref, prefix=case + ".")


def test_edge_attrs_dense_and_lazy(cuda):
    ref = golden("graph_random.npz")
    s_in = ref["bern2.s_in"]
    oracle = go.batch_graph(s_in)
    g_lazy, _ = _build(s_in, cuda)
    g_dense, _ = _build(s_in, cuda, with_edge_attrs=True)
    np.testing.assert_array_equal(g_lazy.edge_attrs.cpu().numpy(), oracle.edge_attrs)
    np.testing.assert_array_equal(g_dense.edge_attrs.cpu().numpy(), oracle.edge_attrs)
    # and back: decode (float type, one-hot) -> uint8 type / dist
    import polyphemus_b200 as pb

    ea = g_lazy.edge_attrs
    t8, d8 = pb.decode_edge_attrs(ea[:, 0], ea[:, 1:])
    assert torch.equal(t8, g_lazy.edge_type) and torch.equal(d8, g_lazy.edge_dist)


def test_single_sequence_matches_reference_semantics(cuda):
    s = go.synthetic_structure(1, 3, 0.2, seed=3)[0]
    s[1] = False                                         # empty bar -> fake activation, self-edge
    oracle = go.sequence_graph(s)
    g, s_dev = _build(s, cuda)
    assert torch.equal(g.batch, g.bars)                  # data.py:202
    np.testing.assert_array_equal(g.edge_index.cpu().numpy(), oracle.edge_index)
    np.testing.assert_array_equal(s_dev.cpu().numpy(), oracle.s_tensor)
    assert bool(s_dev[1, 0, 0])


def test_cpu_input_is_mutated_like_the_reference(cuda):
    import polyphemus_b200 as pb

    s = torch.zeros(2, 2, 4, 32, dtype=torch.bool)
    s[0, 0, 1, 3] = True
    g = pb.graphs_from_tensor(s, device=cuda)
    assert g.num_nodes == 4 and bool(s[0, 1, 0, 0]) and bool(s[1, 0, 0, 0]) and bool(s[1, 1, 0, 0])


@pytest.mark.parametrize("p", [0.1, 0.25, 1.0])
def test_lmd16_full_size_against_oracle_and_invariants(cuda, p):
    """BASELINE config shape (LMD16, batch 256): exact match with the oracle on a slice of sequences plus
    size-independent invariants on the whole batch."""
    bsz = 256
    s_np = go.synthetic_structure(bsz, 16, p, seed=11)
    g, s_dev = _build(s_np, cuda)
    ei = g.edge_index.cpu().numpy()
    et = g.edge_type.cpu().numpy().astype(np.int64)
    ed = g.edge_dist.cpu().numpy().astype(np.int64)
    n = g.num_nodes
    assert n == int(s_dev.sum())
    head = go.batch_graph(s_np[:6])
    e_head = head.edge_index.shape[1]
    np.testing.assert_array_equal(ei[:, :e_head], head.edge_index)
    np.testing.assert_array_equal(et[:e_head], head.edge_type)
    np.testing.assert_array_equal(ed[:e_head], head.edge_dist)
    tail = go.batch_graph(s_np[-3:])
    e_tail = tail.edge_index.shape[1]
    np.testing.assert_array_equal(ei[:, -e_tail:] - (n - tail.num_nodes), tail.edge_index)
    # invariants: edges stay inside a bar, dist range, types, degrees
    gbar = (g.bars + 16 * g.batch).cpu().numpy()
    assert (gbar[ei[0]] == gbar[ei[1]]).all()
    assert ed.min() >= 0 and ed.max() <= 31 and et.min() >= 0 and et.max() <= 5
    assert (ed[et == 4] == 0).all()
    assert np.bincount(ei[1], minlength=n).max() <= 8 and np.bincount(ei[0], minlength=n).max() <= 8
    seg = np.bincount(ei[1] * 6 + et, minlength=n * 6)
    assert seg.max() <= 3 and ((seg.reshape(n, 6) > 0).sum(1) <= 3).all()
    if p == 1.0:
        assert n == bsz * 16 * 128 and ei.shape[1] == bsz * 16 * 1004


def test_csr_plan_matches_edge_list(cuda):
    s_np = go.synthetic_structure(8, 4, 0.3, seed=5)
    g, _ = _build(s_np, cuda)
    plan = g.plan
    ei = g.edge_index.cpu().numpy()
    et = g.edge_type.cpu().numpy().astype(np.int64)
    ed = g.edge_dist.cpu().numpy().astype(np.int64)
    n, e = g.num_nodes, ei.shape[1]
    in_ptr = plan.in_ptr.cpu().numpy()
    in_edge = plan.in_edge.cpu().numpy()
    in_eid = plan.in_eid.cpu().numpy()
    key = ei[1] * 6 + et
    order = np.lexsort((np.arange(e), key))                  # by segment, then by edge id
    np.testing.assert_array_equal(in_eid[:e], order)
    np.testing.assert_array_equal(in_ptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=n * 6))]))
    np.testing.assert_array_equal(in_edge[:e] & 0x3FFFFFF, ei[0][order])
    np.testing.assert_array_equal((in_edge[:e].astype(np.uint32) >> 26).astype(np.int64), ed[order])
    out_ptr = plan.out_ptr.cpu().numpy()
    rec = plan.out_rec.cpu().numpy()
    order_o = np.lexsort((np.arange(e), ei[0]))
    np.testing.assert_array_equal(out_ptr, np.concatenate([[0], np.cumsum(np.bincount(ei[0], minlength=n))]))
    np.testing.assert_array_equal(rec[:e, 0], ei[1][order_o])
    np.testing.assert_array_equal(rec[:e, 1] & 0xFF, et[order_o])
    np.testing.assert_array_equal(rec[:e, 1] >> 8, ed[order_o])
    np.testing.assert_array_equal(rec[:e, 2], order_o)
    seg_len = np.bincount(key, minlength=n * 6)
    np.testing.assert_array_equal(rec[:e, 3], seg_len[key[order_o]])
    # grouping of the out-edge positions by timestep distance (stable) + its work items
    dist_of_pos = ed[order_o]
    perm = plan.dist_perm.cpu().numpy()[:e]
    np.testing.assert_array_equal(perm, np.lexsort((np.arange(e), dist_of_pos)))
    items = plan.dist_items.cpu().numpy()
    item_ptr = plan.dist_item_ptr.cpu().numpy()
    counts = np.bincount(dist_of_pos, minlength=32)
    starts = np.concatenate([[0], np.cumsum(counts)])
    assert item_ptr[0] == 0 and item_ptr[32] <= plan.n_dist_items
    for kk in range(32):
        its = items[item_ptr[kk]:item_ptr[kk + 1]]
        if counts[kk] == 0:
            assert len(its) == 0
            continue
        assert (its[:, 0] == kk).all() and its[0, 1] == starts[kk] and its[-1, 2] == starts[kk + 1]
        assert (its[1:, 1] == its[:-1, 2]).all() and (its[:, 2] > its[:, 1]).all()
    assert (items[item_ptr[32]:, 1] == items[item_ptr[32]:, 2]).all()


def test_graph_build_is_deterministic(cuda):
    s_np = go.synthetic_structure(32, 16, 0.25, seed=2)
    g1, _ = _build(s_np, cuda)
    g2, _ = _build(s_np, cuda)
    assert torch.equal(g1.edge_index, g2.edge_index) and torch.equal(g1.plan.in_eid, g2.plan.in_eid)
    assert torch.equal(g1.plan.out_rec, g2.plan.out_rec)


def test_integration_md_snippet_runs(cuda, built_lib):
    """INTEGRATION.md §2 is the binding a maintainer would copy: execute it verbatim and check it against the oracle."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = [b for b in re.findall(r"```python\n(.*?)```", text, re.S) if "[integration-snippet]" in b]
    assert len(blocks) == 1
    s_np = go.synthetic_structure(5, 3, 0.3, seed=4)
    s_np[2, 1] = False                                        # an empty bar
    env = {"LIB": built_lib, "s_tensor": torch.from_numpy(s_np.copy())}
    exec(compile(blocks[0], "INTEGRATION.md", "exec"), env)
    arrays = go.batch_graph(s_np)
    assert env["N"] == arrays.num_nodes and env["E"] == arrays.edge_index.shape[1]
    np.testing.assert_array_equal(env["edge_index"].cpu().numpy(), arrays.edge_index)
    np.testing.assert_array_equal(env["edge_type"].cpu().numpy(), arrays.edge_type)
    np.testing.assert_array_equal(env["edge_dist"].cpu().numpy(), arrays.edge_dist)
    np.testing.assert_array_equal(env["node_features"].cpu().numpy(), arrays.node_features)
    np.testing.assert_array_equal(env["bars"].cpu().numpy(), arrays.bars)
    ea = env["edge_attrs"].cpu().numpy()
    np.testing.assert_array_equal(ea[:, 0], arrays.edge_type.astype(np.float32))
    np.testing.assert_array_equal(ea[:, 1:].argmax(1), arrays.edge_dist)
    assert env["totals"].numel() == 8
