"""Generate tests/golden/* by executing the reference's OWN unmodified code (container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) — imported in place through oracle.ref_loader on top of the PyG
stand-in (oracle.pyg_shim); nothing from the reference is copied into the repo, only its *outputs* on
seeded synthetic inputs are stored. The fixtures travel to the GPU box, where /root/reference is absent.

Fixtures:
  graph_structure_json.npz  graph_from_tensor(structure.json)            (data.py:141-204)
  graph_random.npz          graph_from_tensor on seeded Bernoulli structures + edge cases, batched
  state_dict_keys.json      names/shapes of VAE(**training.json model) state_dict (255 keys)
  gcl_layer.npz             one GCL fwd + grads (model.py:55-135), dropout 0
  gcn_stack.npz             GCN (2 layers, BatchNorm) fwd + running stats + grads (model.py:190-208)
  vae_step.npz              tiny VAE train step: loss, logits, mu/log_var, all parameter grads
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import graph_oracle as go  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402
from oracle import ref_loader  # noqa: E402


def ref_graph_arrays(ref, s_np):
    """Run the reference graph_from_tensor per sequence + Batch.from_data_list; return plain arrays."""
    from torch_geometric.data import Batch
    s_np = np.array(s_np, dtype=bool, copy=True)
    if s_np.ndim == 3:
        s_np = s_np[None]
    graphs, s_after = [], []
    for b in range(s_np.shape[0]):
        st = torch.from_numpy(s_np[b].copy())
        graphs.append(ref.data.graph_from_tensor(st))
        s_after.append(st.numpy().copy())       # mutated in place (data.py:152-153)
    g = Batch.from_data_list(graphs)
    ea = g.edge_attrs.numpy()
    assert ((ea[:, 1:] == 1).sum(axis=1) == 1).all()
    return dict(
        s_in=s_np, s_out=np.stack(s_after), edge_index=g.edge_index.numpy(),
        edge_type=ea[:, 0].astype(np.int64), edge_dist=ea[:, 1:].argmax(axis=1).astype(np.int64),
        node_features=g.node_features.numpy(), is_drum=g.is_drum.numpy(), bars=g.bars.numpy(),
        batch=g.batch.numpy(), num_nodes=np.int64(int(g.num_nodes)))


def edge_case_structures():
    z = np.zeros((12, 4, 32), dtype=bool)
    # 0: empty bar (fake activation)      1: single node on track 2 (self-edge of type 0)
    z[1, 2, 5] = True
    z[2, :, 7] = True                      # 2: one timestep, all tracks (onset only)
    z[3, 1, :] = True                      # 3: one full track (track edges only)
    z[4] = True                            # 4: full bar, 128 nodes / 1004 edges
    z[5, 0, 0] = z[5, 3, 31] = True        # 5: two nodes, max distance 31 (next edge only)
    z[6, 0, 3] = z[6, 0, 9] = True         # 6: two nodes same track
    z[7, [0, 1], 4] = True                 # 7: two nodes same timestep
    z[8, 0, ::2] = True
    z[8, 1, 1::2] = True                   # 8: alternating tracks -> dense next edges
    z[9, 3, 31] = True                     # 9: single node last cell
    z[10, :, 0] = True
    z[10, :, 31] = True                    # 10: two full timesteps 31 apart
    z[11, 1:, 10:20] = True                # 11: block without drums
    return z


def main():
    ref = ref_loader.load()
    torch.set_num_threads(4)

    # ---------------------------------------------------------------- graphs
    s_json = np.array(json.load(open(os.path.join(ref_loader.REFERENCE_DIR, "structure.json"))), dtype=bool)
    np.savez_compressed(os.path.join(HERE, "graph_structure_json.npz"), **ref_graph_arrays(ref, s_json))

    rng = np.random.default_rng(20261017)
    cases = {}
    for i, p in enumerate((0.02, 0.1, 0.25, 0.5, 0.9)):
        cases[f"bern{i}"] = rng.random((3, 4, 4, 32)) < p          # [B=3, n_bars=4, 4, 32]
    cases["edge"] = edge_case_structures().reshape(3, 4, 4, 32)
    cases["lmd16"] = rng.random((2, 16, 4, 32)) < 0.25
    out = {}
    for name, s in cases.items():
        for k, v in ref_graph_arrays(ref, s).items():
            out[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(HERE, "graph_random.npz"), **out)

    # ---------------------------------------------------------------- state dict contract
    cfg = json.load(open(os.path.join(ref_loader.REFERENCE_DIR, "training.json")))["model"]
    torch.manual_seed(0)
    vae = ref.model.VAE(**cfg, device=torch.device("cpu"))
    keys = {k: list(v.shape) for k, v in vae.state_dict().items()}
    n_params = sum(p.numel() for p in vae.parameters())
    json.dump({"config": cfg, "n_params": n_params, "keys": keys},
              open(os.path.join(HERE, "state_dict_keys.json"), "w"), indent=0)
    del vae

    # ---------------------------------------------------------------- one GCL layer
    d = 64
    torch.manual_seed(1)
    arrays = go.batch_graph(go.synthetic_structure(2, 2, 0.3, seed=5))
    ei = torch.from_numpy(arrays.edge_index)
    ea = torch.from_numpy(arrays.edge_attrs)
    edge_nn = torch.nn.Linear(32, d)
    layer = ref.model.GCL(d, d, 6, edge_nn, dropout=0.0)
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    x = torch.randn(arrays.num_nodes, d, requires_grad=True)
    y = layer(x, ei, ea[:, 0], ea[:, 1:])
    gy = torch.randn_like(y)
    y.backward(gy)
    np.savez_compressed(
        os.path.join(HERE, "gcl_layer.npz"), x=x.detach().numpy(), edge_index=arrays.edge_index,
        edge_type=arrays.edge_type, edge_dist=arrays.edge_dist, weight=layer.weight.detach().numpy(),
        root=layer.root.detach().numpy(), bias=layer.bias.detach().numpy(),
        nn_weight=edge_nn.weight.detach().numpy(), nn_bias=edge_nn.bias.detach().numpy(),
        y=y.detach().numpy(), gy=gy.numpy(), gx=x.grad.numpy(), g_weight=layer.weight.grad.numpy(),
        g_root=layer.root.grad.numpy(), g_bias=layer.bias.grad.numpy(),
        g_nn_weight=edge_nn.weight.grad.numpy(), g_nn_bias=edge_nn.bias.grad.numpy())

    # ---------------------------------------------------------------- GCN stack (BN, residual)
    torch.manual_seed(2)
    gcn = ref.model.GCN(input_dim=d, hidden_dim=d, n_layers=2, num_relations=6, batch_norm=True, dropout=0)
    for lyr in gcn.layers:
        lyr.dropout = 0.0
    with torch.no_grad():
        for p in gcn.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    gcn.train()
    sd0 = {k: v.detach().clone().numpy() for k, v in gcn.state_dict().items()}
    x = torch.randn(arrays.num_nodes, d, requires_grad=True)
    data = type("D", (), {})()
    data.x, data.edge_index, data.edge_attrs = x, ei, ea
    y = gcn(data)
    gy = torch.randn_like(y)
    y.backward(gy)
    gcn_out = {f"sd.{k}": v for k, v in sd0.items()}
    gcn_out.update({f"sd_after.{k}": v.detach().numpy() for k, v in gcn.state_dict().items() if "running" in k})
    gcn_out.update({f"grad.{k}": p.grad.numpy() for k, p in gcn.named_parameters()})
    np.savez_compressed(os.path.join(HERE, "gcn_stack.npz"), x=x.detach().numpy(), y=y.detach().numpy(),
                        gy=gy.numpy(), gx=x.grad.numpy(), edge_index=arrays.edge_index,
                        edge_type=arrays.edge_type, edge_dist=arrays.edge_dist, **gcn_out)
    # eval-mode forward (running stats)
    gcn.eval()
    with torch.no_grad():
        data.x = x.detach()
        y_eval = gcn(data)
    np.savez_compressed(os.path.join(HERE, "gcn_stack_eval.npz"), y_eval=y_eval.numpy())

    # ---------------------------------------------------------------- tiny VAE training step
    from torch_geometric.data import Batch
    cfg_small = dict(dropout=0, batch_norm=True, gnn_n_layers=2, d=64, n_bars=2, resolution=8)
    torch.manual_seed(3)
    vae = ref.model.VAE(**cfg_small, device=torch.device("cpu"))
    for m in vae.modules():
        if isinstance(m, ref.model.GCL):
            m.dropout = 0.0
    with torch.no_grad():
        for p in vae.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    vae.train()
    bsz = 4
    s_np = go.synthetic_structure(bsz, 2, 0.25, seed=7)
    s_np[1, 1] = False                                   # one empty bar -> fake activation path
    arrays = go.batch_graph(s_np)
    tokens = mo.synthetic_tokens(arrays.num_nodes, seed=7)
    c_all = mo.onehot_content(tokens)
    graphs, off = [], 0
    for b in range(bsz):
        st = torch.from_numpy(s_np[b].copy())
        g = ref.data.graph_from_tensor(st)
        g.c_tensor = c_all[off:off + int(g.num_nodes)]
        off += int(g.num_nodes)
        g.s_tensor = st.float()
        graphs.append(g)
    batch = Batch.from_data_list(graphs)
    sd0 = {k: v.detach().clone().numpy() for k, v in vae.state_dict().items()}
    noise = torch.randn(bsz, cfg_small["d"], generator=torch.Generator().manual_seed(11))
    orig_randn_like = torch.randn_like
    torch.randn_like = lambda t, *a, **k: noise.to(t.dtype)      # pin the reparameterisation noise
    try:
        (s_logits, c_logits), mu, log_var = vae(batch)
    finally:
        torch.randn_like = orig_randn_like
    trainer = ref.training.PolyphemusTrainer("/tmp/none", vae, None)
    trainer.beta = 0                                              # training.py:116
    tot, parts = trainer._losses(batch.s_tensor, s_logits, batch.c_tensor, c_logits, mu, log_var)
    tot.backward()
    step = {f"sd.{k}": v for k, v in sd0.items()}
    for k, p in vae.named_parameters():
        step[f"grad.{k}"] = (p.grad.numpy() if p.grad is not None else np.zeros(0, dtype=np.float32))
    step.update({f"sd_after.{k}": v.detach().numpy() for k, v in vae.state_dict().items() if "running" in k})
    np.savez_compressed(
        os.path.join(HERE, "vae_step.npz"), s_in=s_np, tokens=tokens.numpy().astype(np.int16),
        noise=noise.numpy(), s_logits=s_logits.detach().numpy(), c_logits=c_logits.detach().numpy(),
        mu=mu.detach().numpy(), log_var=log_var.detach().numpy(), loss=np.float32(float(tot)),
        loss_parts=np.array([parts[k] for k in ("pitch", "dur", "structure", "kld")], dtype=np.float32),
        config=json.dumps(cfg_small), **step)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(HERE, f)):>9d} B")


if __name__ == "__main__":
    main()
