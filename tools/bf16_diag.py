"""Diagnostic: per-parameter gradient cosine / norm ratio of the bf16 bench mode against the fp32 oracle (LMD2, B=64)."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import polyphemus_b200 as pb
import polyphemus_b200.ops as ops_mod
from polyphemus_b200 import ops
import test_parity_scale_gpu as T

cuda = torch.device("cuda", 0)
for label, p_gcl, act_bf16, autocast in (("bench mode", 0.1, True, True), ("no dropout", 0.0, True, True),
                                         ("fp32 act storage", 0.0, False, True), ("fp32 act, no autocast", 0.0, False, False)):
    ops.set_bf16_activations(act_bf16)
    vae, cfg, sd_cpu, arrays, tokens, noise, graph = T._setup(cuda, 2, 64, "bf16", p_gcl, seed=300)
    n_layers, d, n_edges = cfg["gnn_n_layers"], cfg["d"], graph.num_edges
    torch.manual_seed(4242)
    ops_mod._seed_counter = itertools.count()
    masks = None
    if p_gcl > 0:
        seeds = [ops.next_seed() for _ in range(2 * n_layers)]
        masks = {}
        for i, s in enumerate(seeds):
            prefix = "encoder.c_encoder.graph_encoder" if i < n_layers else "decoder.c_decoder.graph_decoder"
            masks[(prefix, i % n_layers)] = ops.dropout_keep_mask(n_edges, d, p_gcl, s, cuda).cpu()
        ops_mod._seed_counter = itertools.count()
    loss, _ = T._train_step(vae, graph, noise, cuda, bf16=autocast)
    sd, _, loss_ref, _ = T._oracle_step(cfg, sd_cpu, arrays, tokens, noise, keep_masks=masks, p_gcl=p_gcl)
    print(f"== {label}: loss {loss:.6f} vs oracle {loss_ref:.6f}")
    rows = []
    for name, p in vae.named_parameters():
        want = sd[name].grad
        if want is None or T._zero_grad_by_math(name):
            continue
        a, b = p.grad.detach().double().cpu().flatten(), want.double().flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp(min=1e-300))
        rows.append((name, cos, float(a.norm() / b.norm().clamp(min=1e-300)), float(b.norm())))
    for name, cos, ratio, nb in rows:
        if label == "bench mode" or cos < 0.999:
            print(f"   {name:70s} cos {cos:.5f} ratio {ratio:.4f} |g| {nb:.3e}")
    print("   min cos", min(r[1] for r in rows))
pb.set_precision("fp32")
