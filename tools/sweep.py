"""Secondary BASELINE.json configurations on one B200 (not the bench line; evidence for profiles/):

  layer    isolated relational graph-conv layer (GCL + BN + ReLU + residual) fwd+bwd, E in {1e5, 1e6, 1e7},
           d in {256, 512, 1024}, bf16 and fp32(tf32x3) — per-kernel CUDA-event times vs the HBM / tensor roofline
  batch    LMD16 training step at per-GPU batch 256..2048 (seq/s, peak memory)
  generate LMD2 decoder-only generation of 4096 sequences with structure.json conditioning (and unconditioned)
  density  LMD16 batch 256 at structure densities 0.1 / 0.25 / 0.5 / 1.0 (graph size, build time, step time)
  precision LMD16 batch 256 step in the fp32-grade mode next to bf16

    python tools/sweep.py layer|batch|generate  > gpurun_out/sweep_<name>.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import polyphemus_b200 as pb
from polyphemus_b200 import _ffi
from polyphemus_b200.train import TrainStep, device_batch, synthetic_host_batch

dev = torch.device("cuda", 0)
PK = bench.peaks()


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def layer_sweep():
    for d in (256, 512, 1024):
        for target_e in (1e5, 1e6, 1e7):
            n_bars = max(1, int(target_e / 112))                    # ~112 edges per bar at p = 0.25
            if n_bars * 32 * d * 4 * 30 > 120e9:                    # activations + operands would not fit comfortably
                continue
            rng = np.random.default_rng(0)
            s = torch.from_numpy(rng.random((n_bars, 1, 4, 32)) < 0.25).to(dev)
            graph = pb.graphs_from_tensor(s)
            n, e = graph.num_nodes, graph.num_edges
            for precision in ("bf16", "fp32"):
                if precision == "fp32" and n * 7 * d * 8 * 3 > 100e9:
                    continue
                gcn = pb.GCN(input_dim=d, hidden_dim=d, n_layers=1, num_relations=6, batch_norm=True, dropout=0,
                             precision=precision).to(dev).train()
                x = torch.randn(n, d, device=dev, requires_grad=True)
                gy = torch.randn(n, d, device=dev)

                def step():
                    graph.x = x
                    y = gcn(graph)
                    y.backward(gy)
                    x.grad = None

                for _ in range(3):
                    step()
                _ffi.profiler.reset()
                _ffi.profiler.enabled = True
                iters = 5
                ms = timed(step, iters)
                _ffi.profiler.enabled = False
                summ = _ffi.profiler.summary()
                st = graph.structured if pb.ops.structured_enabled() and d % 256 == 0 else None
                rows = bench.kernel_table(summ, n, e, d, iters, precision, PK, 0.1, slots=3 if st else 6,
                                          n_rows=st.n_padded if st else n,
                                          act_bytes=2 if (precision == "bf16" and st and pb.ops.bf16_activations_enabled()) else 4)
                print(json.dumps({"config": "layer", "d": d, "nodes": n, "edges": e, "precision": precision,
                                  "ms_fwd_bwd": ms, "layout": "structured" if st else "generic",
                                  "kernels": [{k: r[k] for k in ("kernel", "bound", "achieved", "unit", "frac", "avg_ms")}
                                              for r in rows if r["frac"] is not None]}), flush=True)
                del gcn, x, gy
            del graph
            torch.cuda.empty_cache()


def batch_sweep():
    pb.set_precision("bf16")
    for batch in (256, 512, 1024, 2048):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        try:
            torch.manual_seed(0)
            model = pb.VAE(**bench.MODEL_CFG, device=dev).to(dev).train()
            step = TrainStep(model, autocast_bf16=True, **bench.ADAM)
            host = synthetic_host_batch(batch, 16, 0.25, seed=0)
            fn = lambda: step(device_batch(host, dev))
            for _ in range(3):
                fn()
            ms = timed(fn, 5)
            print(json.dumps({"config": "batch", "per_gpu_batch": batch, "ms_per_step": ms, "seq_per_s": batch / ms * 1e3,
                              "nodes": int(host.tokens.size(0)), "peak_mem_gib": torch.cuda.max_memory_allocated() / 2**30}),
                  flush=True)
            del model, step
        except torch.OutOfMemoryError as exc:
            print(json.dumps({"config": "batch", "per_gpu_batch": batch, "error": "out of memory", "detail": str(exc)[:120]}),
                  flush=True)


def generate_sweep():
    cfg = dict(bench.MODEL_CFG, n_bars=2)
    torch.manual_seed(0)
    vae = pb.VAE(**cfg, device=dev).to(dev).eval()
    s_json = torch.from_numpy(np.load(os.path.join(bench.ROOT, "tests", "golden", "graph_structure_json.npz"))["s_in"][0]).bool()
    n = 4096
    for precision in ("bf16", "fp32"):
        pb.set_precision(precision)
        for mode in ("conditioned", "unconditioned"):
            def gen():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
                    z = torch.randn(n, cfg["d"], device=dev)
                    s = vae.decoder._structure_from_binary(s_json.unsqueeze(0).repeat(n, 1, 1, 1).to(dev)) if mode == "conditioned" else None
                    s_logits, c_logits = vae.decoder(z, s)
                    return c_logits
            for _ in range(2):
                out = gen()
            ms = timed(gen, 5)
            print(json.dumps({"config": "generate", "sequences": n, "mode": mode, "precision": precision, "ms": ms,
                              "seq_per_s": n / ms * 1e3, "nodes": int(out.shape[0])}), flush=True)


def density_sweep():
    """LMD16 batch 256 at structure densities 0.1 .. 1.0: graph size, device graph build time, bf16 step time."""
    pb.set_precision("bf16")
    for p in (0.1, 0.25, 0.5, 1.0):
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        try:
            host = synthetic_host_batch(256, 16, p, seed=0)
            build_ms = timed(lambda: device_batch(host, dev), 5)
            graph = device_batch(host, dev)
            model = pb.VAE(**bench.MODEL_CFG, device=dev).to(dev).train()
            step = TrainStep(model, autocast_bf16=True, **bench.ADAM)
            fn = lambda: step(device_batch(host, dev))
            for _ in range(3):
                fn()
            ms = timed(fn, 5)
            print(json.dumps({"config": "density", "p": p, "nodes": graph.num_nodes, "edges": graph.num_edges,
                              "edges_per_node": graph.num_edges / graph.num_nodes, "graph_build_ms": build_ms,
                              "ms_per_step": ms, "seq_per_s": 256 / ms * 1e3,
                              "peak_mem_gib": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
            del model, step, graph
        except torch.OutOfMemoryError as exc:
            print(json.dumps({"config": "density", "p": p, "error": "out of memory", "detail": str(exc)[:120]}), flush=True)


def precision_sweep():
    """LMD16 batch 256 step in the fp32-grade mode (TF32x3 tensor-core split) next to bf16."""
    for precision in ("bf16", "fp32"):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        pb.set_precision(precision)
        torch.manual_seed(0)
        model = pb.VAE(**bench.MODEL_CFG, device=dev).to(dev).train()
        step = TrainStep(model, autocast_bf16=precision == "bf16", **bench.ADAM)
        host = synthetic_host_batch(256, 16, 0.25, seed=0)
        fn = lambda: step(device_batch(host, dev))
        for _ in range(3):
            fn()
        ms = timed(fn, 5)
        print(json.dumps({"config": "precision", "precision": precision, "ms_per_step": ms, "seq_per_s": 256 / ms * 1e3,
                          "peak_mem_gib": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
        del model, step
    pb.set_precision("bf16")


if __name__ == "__main__":
    {"layer": layer_sweep, "batch": batch_sweep, "generate": generate_sweep, "density": density_sweep,
     "precision": precision_sweep}[sys.argv[1]]()
