#!/usr/bin/env python
"""bench.py — LMD16 VAE training step (graph build + fwd + loss + bwd + gradient all-reduce + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision bf16|fp32] [--impl reference]

Contract (see the task statement): prints ONE JSON line. `value` = sequences/s with the step's inputs (bool
structure tensor + int16 note tokens) already resident in HBM; `e2e` = the same step fed from pinned HOST
buffers through the public API (H2D copies and a D2H read of the loss inside the timed region); `roofline` =
the dominant kernel of the message-passing path timed live with CUDA events over the timed region;
`cpu_baseline` = the oracle port (oracle/model_oracle.py) on this box's host cores, bounded sample.
For N > 1 launch with torchrun (one rank per GPU, NCCL); batch per GPU is fixed (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL_CFG = dict(dropout=0, batch_norm=True, gnn_n_layers=8, d=512, n_bars=16, resolution=8)  # training.json + n_bars=16
ADAM = dict(lr=5e-6, betas=(0.9, 0.98), eps=1e-9)                                             # training.json optimizer
DENSITY = 0.25
WORKLOAD = "LMD16 VAE train step (16 bars x 4 tracks x 32 steps), synthetic Bernoulli(0.25) pianorolls, random init"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.split(",") for r in open(self.file.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.file.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_step_factory(batch: int, n_bars: int, gcl_dropout: float, seed: int = 0):
    """One graph build + fwd + loss + bwd of the oracle port (the reference's algorithm in plain PyTorch on the CPU:
    per-relation boolean compaction, Linear on one-hot distances, index_add scatter-mean, seven matmuls per layer,
    model.py:103-135) on a bounded sample. The GCL message dropout is drawn exactly as the reference draws it
    (F.dropout on every relation's E_r x d message tensor, model.py:133) — the same workload as the GPU arm."""
    from oracle import graph_oracle as go
    from oracle import model_oracle as mo
    import polyphemus_b200 as pb

    cfg = dict(MODEL_CFG, n_bars=n_bars)
    torch.manual_seed(0)
    sd0 = pb.VAE(**cfg, device=torch.device("cpu")).state_dict()   # parameter container only (random init)
    sd = mo.leaf_state(sd0)
    s_np = go.synthetic_structure(batch, n_bars, DENSITY, seed)

    def step():
        arrays = go.batch_graph(s_np)                               # graph construction is part of the step
        tokens = mo.synthetic_tokens(arrays.num_nodes, seed)
        gb = mo.make_batch(arrays, tokens)
        ctx = mo.Ctx(training=True, gcl_dropout=gcl_dropout, gcl_random_dropout=gcl_dropout > 0)
        (s_logits, c_logits), mu, log_var = mo.vae(sd, gb, n_bars, cfg["d"], ctx)
        loss, _ = mo.losses(gb.s_tensor, s_logits, gb.c_tensor, c_logits, mu, log_var)
        for v in sd.values():
            if v.grad is not None:
                v.grad = None
        loss.backward()
        return float(loss.detach())

    return step


def time_cpu_baseline(batch: int, n_bars: int, steps: int, warmup: int, gcl_dropout: float):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_oracle_step_factory(batch, n_bars, gcl_dropout)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return batch / min(times), times


def cpu_graph_build_ms_per_bar(n_bars: int = 16, seqs: int = 4) -> float:
    """data.graph_from_tensor's algorithm (oracle.graph_oracle, numpy) on one core: ms per bar."""
    from oracle import graph_oracle as go

    torch.set_num_threads(1)
    s_np = go.synthetic_structure(seqs, n_bars, DENSITY, 0)
    go.batch_graph(s_np[:1])
    t0 = time.perf_counter()
    go.batch_graph(s_np)
    dt = time.perf_counter() - t0
    torch.set_num_threads(os.cpu_count() or 1)
    return 1e3 * dt / (seqs * n_bars)


def cpu_baseline_block(args) -> dict:
    """LMD16 batch 8 (the metric's sequence shape), config 1 exactly (LMD2, batch 64), graph
    construction ms/bar — all on this box's host cores, GCL dropout as in the GPU arm."""
    v16, t16 = time_cpu_baseline(args.cpu_batch, MODEL_CFG["n_bars"], 2, 1, args.gcl_dropout)
    v2, t2 = time_cpu_baseline(64, 2, 2, 1, args.gcl_dropout)
    return {"value": v16, "unit": "seq/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle port (reference algorithm, PyTorch CPU fp32, GCL dropout {args.gcl_dropout} drawn as "
                      f"model.py:133), LMD16 batch {args.cpu_batch}: graph build + fwd + loss + bwd, best of {len(t16)} "
                      f"after 1 warm-up ({min(t16):.2f} s/step)",
            "config1_lmd2_batch64": {"value": v2, "unit": "seq/s", "s_per_step": min(t2),
                                     "sample": "BASELINE.json configs[0]: LMD2, batch 64, fwd + loss + bwd, best of 2 after 1 warm-up"},
            "graph_build_ms_per_bar": cpu_graph_build_ms_per_bar(),
            "same_config_as_gpu_arm": {"model": True, "gcl_dropout": True, "batch": False},
            "note": "the port is ~7x faster than the reference's own files measured at survey time (BASELINE.md §2: "
                    "1.4 seq/s LMD16 batch 8, 12.9 seq/s LMD2 batch 64, 3.3-4.1 ms/bar on 8 vCPU)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    t_all = time.perf_counter()
    value, times = time_cpu_baseline(batch, MODEL_CFG["n_bars"], max(1, args.steps), max(0, min(args.warmup, 1)),
                                     args.gcl_dropout)
    ms = 1e3 * statistics.mean(times)
    sample = (f"oracle port (reference algorithm, PyTorch CPU fp32, GCL dropout {args.gcl_dropout}) of the LMD16 step on "
              f"batch {batch} (graph build + fwd + loss + bwd), best of {len(times)}")
    line = {
        "impl": "reference", "metric": "LMD16 train seqs/s", "value": value, "unit": "seq/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch": batch, "n_bars": 16, "d": 512, "gnn_n_layers": 8,
                   "gcl_dropout": args.gcl_dropout},
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ roofline
def ncu_traffic() -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel family, from the committed `ncu --set full`
    captures (profiles/traffic.json, written by tools/summarize_profiles.py) — measured once under the profiler, not
    during this run."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as fh:
        return json.load(fh)


def kernel_table(summary: dict, n: int, e: int, d: int, steps: int, precision: str, pk: dict, p_drop: float,
                 slots: int = 6, n_rows: int = 0, act_bytes: int = 4):
    """Per-kernel-family algorithmic bytes / flops per launch and achieved rates.
    `slots` = operand blocks per node (6 relations, or 3 in the structured layout), `n_rows` = GEMM rows (padded)."""
    r = slots
    k = (r + 1) * d
    n_rows = n_rows or n
    s = 2 if precision == "bf16" else 8        # bytes per GEMM-operand element (bf16, or TF32 hi+lo fp32 pair)
    s_da = 2 if precision == "bf16" else 4
    eb = 8 if p_drop > 0 else 4
    ab = act_bytes                          # activation storage between the kernels of a stack (x, out, y, gy, gx)
    parts = 1184 * d * 4                    # per-item partial rows of the edge-table gradient
    algo = {
        "pb_agg_fwd": ("hbm", n * d * ab + eb * e + 4 * (n * r + 1) + 128 * d + n * k * s),
        "pb_agg_bwd": ("hbm", n * k * s_da + 2 * n * d * ab + 16 * e + 4 * (n + 1) + n * d * ab + 128 * d + parts),
        # the fused backward moves the same compulsory data (its work list is 16 B per edge + 48 B per node instead of
        # 16 B per edge + 4 B per node; the denominator is kept identical to round 1's for comparability)
        "pb_agg_bwd_fused": ("hbm", n * k * s_da + 2 * n * d * ab + 16 * e + 4 * (n + 1) + n * d * ab + 128 * d + parts),
        "pb_bn_stats": ("hbm", n * d * ab),
        "pb_bn_relu_res_fwd": ("hbm", 3 * n * d * ab),
        "pb_bn_relu_res_bwd": ("hbm", 4 * n * d * ab + n * d * s),
        "pb_rgcn_gemm_fwd": ("tensor", 2 * n_rows * k * d),          # executed flops (structured: 4d-wide operand)
        "pb_rgcn_gemm_fwd_bn": ("tensor", 2 * n_rows * k * d),       # the same GEMM, BatchNorm column sums in its epilogue
        "pb_rgcn_gemm_bwd_data": ("tensor", 2 * n_rows * k * d),
        "pb_rgcn_gemm_bwd_weight": ("tensor", 2 * n_rows * k * d),
    }
    mma_passes = 1 if precision == "bf16" else 3
    rows = []
    for name, (bound, work) in algo.items():
        if name not in summary:
            continue
        ent = summary[name]
        # `work` is per ABI call = per layer (a call may be several launches, e.g. the grouped weight gradient; an
        # operand recompute in backward is one more pb_agg_fwd call)
        per_layer_s = ent["ms"] / ent["calls"] * 1e-3
        if bound == "hbm":
            achieved, peak, unit = work / per_layer_s / 1e9, pk["hbm"], "GB/s"
        else:
            achieved, peak, unit = work / per_layer_s / 1e12, pk["tf_sustained"], "TFLOP/s"
        rows.append({"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                     "frac": achieved / peak, "avg_ms": per_layer_s * 1e3, "launches_per_step": ent["calls"] / steps,
                     "ms_per_step": ent["ms"] / steps, "algorithmic_per_launch": work,
                     **({"executed_mma_passes": mma_passes,
                         "dense_reference_flops_per_launch": 2 * n * 7 * d * d} if bound == "tensor" else {})})
    for name, ent in summary.items():           # dense layers next to the path, reported without a roofline claim
        if name.endswith(":linear"):
            rows.append({"kernel": name, "bound": "tensor", "achieved": None, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": None, "avg_ms": ent["ms"] / ent["calls"], "launches_per_step": ent["calls"] / steps,
                         "ms_per_step": ent["ms"] / steps, "algorithmic_per_launch": None})
    rows.sort(key=lambda x: -x["ms_per_step"])
    return rows


# ------------------------------------------------------------------------------------------------ secondary configs
def _timed_ms(fn, iters, world=1):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def _train_ms(batch, precision, dev, world, rank, args, warm=3, iters=4, prebuilt=False):
    """ms/step of the LMD16 training step at `batch` sequences per GPU (device graph build included, no prefetch).
    `prebuilt`: graphs and their CSR plans built before the timed region (the step with graph construction
    excluded, as behind the reference's DataLoader workers)."""
    import polyphemus_b200 as pb
    from polyphemus_b200.train import TrainStep, device_batch, synthetic_host_batch

    pb.set_precision(precision)
    torch.manual_seed(0)
    model = pb.VAE(**MODEL_CFG, device=dev).to(dev).train()
    for m in model.modules():
        if isinstance(m, pb.GCL):
            m.dropout = args.gcl_dropout
    step = TrainStep(model, autocast_bf16=precision == "bf16", **ADAM)
    host = synthetic_host_batch(batch, MODEL_CFG["n_bars"], DENSITY, seed=7 + rank)
    try:
        if prebuilt:
            graphs = [device_batch(host, dev) for _ in range(3)]
            for g in graphs:
                g.structured                                       # CSR plans of the structured layout
            it = iter(range(10 ** 9))
            fn = lambda: step(graphs[next(it) % len(graphs)])
        else:
            fn = lambda: step(device_batch(host, dev))
        for _ in range(warm):
            fn()
        ms = _timed_ms(fn, iters, world)
        return {"per_gpu_batch": batch, "global_batch": batch * world, "ms_per_step": ms, "seq_per_s": batch * world / ms * 1e3,
                "nodes_per_gpu": int(host.tokens.size(0)), "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2**30}
    finally:
        step.close()
        pb.set_precision(args.precision)


def _layer_point(target_e, d, precision, dev, pk, p_drop=0.1):
    """One GCN layer (GCL + BatchNorm + ReLU + residual, model.py:198-206) fwd + bwd in isolation on an LMD-shaped
    graph with ~target_e edges, called on the stack's own (structured, padded) layout so that only the layer's
    kernels run."""
    import polyphemus_b200 as pb
    from polyphemus_b200 import _ffi

    n_bars = max(16, int(target_e / 112) // 16 * 16)              # ~112 edges per bar at p = 0.25
    gen = torch.Generator(device=dev).manual_seed(0)
    s = torch.rand((n_bars // 16, 16, 4, 32), device=dev, generator=gen) < DENSITY
    graph = pb.graphs_from_tensor(s)
    del s
    n, e = graph.num_nodes, graph.num_edges
    st = graph.structured
    bf16 = precision == "bf16"
    gcn = pb.GCN(input_dim=d, hidden_dim=d, n_layers=1, num_relations=6, batch_norm=True, dropout=0, precision=precision).to(dev).train()
    layer, bn = gcn.layers[0], gcn.norm_layers[0].module
    layer.dropout = p_drop
    adt = torch.bfloat16 if bf16 else torch.float32
    x = torch.randn(st.n_padded, d, device=dev, dtype=adt).requires_grad_(True)
    gy = torch.randn(st.n_padded, d, device=dev, dtype=adt)

    def step():
        layer(x, plan=st.plan, bn=bn, struct=st).backward(gy)
        x.grad = None

    for _ in range(2):
        step()
    _ffi.profiler.reset()
    _ffi.profiler.enabled = True
    iters = 3
    ms = _timed_ms(step, iters)
    _ffi.profiler.enabled = False
    summ = _ffi.profiler.summary()
    rows = kernel_table(summ, n, e, d, iters, precision, pk, p_drop, slots=3, n_rows=st.n_padded, act_bytes=2 if bf16 else 4)
    hb = [r for r in rows if r["bound"] == "hbm"]
    hbm_ms = sum(r["ms_per_step"] for r in hb)
    hbm_bytes = sum(r["algorithmic_per_launch"] * r["launches_per_step"] for r in hb)
    return {"edges": e, "nodes": n, "d": d, "precision": precision, "gcl_dropout": p_drop, "ms_fwd_bwd": ms,
            "hbm_kernels_frac": hbm_bytes / (hbm_ms * 1e-3) / 1e9 / pk["hbm"] if hbm_ms else None,
            "kernels": {r["kernel"]: round(r["frac"], 3) for r in rows if r["frac"] is not None},
            "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2**30}


def secondary_measurements(args, dev, world, rank, pk):
    """The other BASELINE.json configurations, short runs (a few steps each) after the headline measurement:
    configs[1] fp32 arm, configs[2] LMD2 generation of 4096 sequences to the pianoroll tensor, configs[3] isolated
    layer sweep up to 1e8 edges (N = 1 only), configs[4] per-GPU batch 512 .. 2048 at this N. Failures (e.g. out of
    memory at the largest points) are recorded, never fatal."""
    import gc
    import numpy as np
    import torch.distributed as dist
    import polyphemus_b200 as pb

    out = {}

    def guarded(name, fn):
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        ok = 1
        try:
            res = fn()
        except Exception as exc:   # noqa: BLE001 - recorded in the JSON line
            res, ok = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}, 0
        if world > 1:              # every rank must agree before the next collective-bearing measurement
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0 and ok:
                res = {"error": "another rank failed"}
        out[name] = res
        return ok

    # configs[4]: large per-GPU batches at this N (weak scaling, NCCL all-reduce inside the step)
    for b in (512, 1024, 2048):
        gc.collect()
        torch.cuda.empty_cache()
        free = torch.cuda.mem_get_info(dev)[0] / 2**30
        need = {512: 50, 1024: 85, 2048: 150}[b]
        enough = torch.tensor([1 if free > need else 0], device=dev)
        if world > 1:
            dist.all_reduce(enough, op=dist.ReduceOp.MIN)
        if int(enough) == 0:
            out[f"large_batch_{b}"] = {"skipped": f"{free:.0f} GiB free < {need} GiB needed"}
            continue
        if not guarded(f"large_batch_{b}", lambda: _train_ms(b, "bf16", dev, world, rank, args, warm=2, iters=3)):
            break
    if world == 1:
        # configs[1], fp32 arm: the parity mode (TF32x3 on the tensor cores) on the headline workload
        guarded("fp32_mode_batch256", lambda: _train_ms(args.batch, "fp32", dev, 1, rank, args, warm=2, iters=3))
        # the headline step with graph construction excluded (graphs + plans prebuilt)
        guarded("prebuilt_graphs_batch256", lambda: _train_ms(args.batch, "bf16", dev, 1, rank, args, warm=3, iters=6,
                                                              prebuilt=True))

        # configs[2]: LMD2 decoder-only generation of 4096 sequences with structure conditioning, down to the
        # [4096, 2, 4, 32, 15, 230] pianoroll tensor (generate.py:24-35, 226-237, utils.py:59-79)
        def gen():
            cfg = dict(MODEL_CFG, n_bars=2)
            pb.set_precision("bf16")
            torch.manual_seed(0)
            vae = pb.VAE(**cfg, device=dev).to(dev).eval()
            s_json = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "graph_structure_json.npz"))["s_in"][0]).bool()
            n = 4096
            s_cond = s_json.unsqueeze(0).repeat(n, 1, 1, 1).to(dev)

            def run():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    z = torch.randn(n, cfg["d"], device=dev)
                    graph = vae.decoder._structure_from_binary(s_cond.clone())
                    _, c_logits = vae.decoder(z, graph)
                    return pb.mtp_from_logits(c_logits.to(torch.bfloat16), s_cond)

            for _ in range(2):
                mtp = run()
            ms = _timed_ms(run, 3)
            res = {"sequences": n, "ms": ms, "seq_per_s": n / ms * 1e3, "nodes": int(s_cond.sum()),
                   "mtp_shape": list(mtp.shape), "mtp_dtype": "bf16", "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2**30}
            pb.set_precision(args.precision)
            return res
        guarded("generation_lmd2_4096", gen)

        # configs[3]: isolated layer sweep
        for target_e, d, precision, p_drop in ((1e5, 512, "bf16", 0.1), (1e6, 256, "bf16", 0.1), (1e6, 512, "bf16", 0.1),
                                               (1e6, 1024, "bf16", 0.1), (1e6, 512, "fp32", 0.1), (1e7, 512, "bf16", 0.1),
                                               (1e7, 1024, "bf16", 0.1), (3e7, 256, "bf16", 0.1), (1e8, 256, "bf16", 0.0)):
            guarded(f"layer_E{target_e:.0e}_d{d}_{precision}".replace("+0", ""),
                    lambda: _layer_point(target_e, d, precision, dev, pk, p_drop))
    pb.set_precision(args.precision)
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU (training.json batch_size)")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-batch", type=int, default=8, help="bounded CPU sample (sequences)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary BASELINE.json configurations")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: skip the second (host-fed) timed region")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the first timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--gcl-dropout", type=float, default=0.1, help="GCL message dropout (hard-wired 0.1 in the reference)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import polyphemus_b200 as pb
    from polyphemus_b200 import _ffi
    from polyphemus_b200.train import TrainStep, device_batch, synthetic_host_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: polyphemus_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    pb.set_precision(args.precision)
    torch.manual_seed(0)
    model = pb.VAE(**MODEL_CFG, device=dev).to(dev).train()
    for m in model.modules():
        if isinstance(m, pb.GCL):
            m.dropout = args.gcl_dropout
    step_fn = TrainStep(model, autocast_bf16=args.precision == "bf16", **ADAM)

    # a few distinct batches so the step cannot specialise on one graph; each rank gets its own shard
    n_variants = 2
    # e2e input = the samples exactly as they lie on disk (preprocess.py:210: int16 c_tensor [4, T, 16, 2] + bool s_tensor
    # [4, T] per sample), pinned; bars reshape / fake activations / silence filter / graph build happen on the device
    hosts = [synthetic_host_batch(args.batch, MODEL_CFG["n_bars"], DENSITY, seed=1000 * rank + i, disk=True)
             for i in range(n_variants)]
    resident = [(h.s_tensor.to(dev), h.tokens.to(dev)) for h in hosts]

    # Every step builds its graph on the device; the build of step i+1 is issued on a side stream before step i is
    # enqueued (train.BatchPrefetcher), as a data loader would, so its size read-back does not drain the training stream.
    from polyphemus_b200.train import BatchPrefetcher, HostBatch
    prefetch = BatchPrefetcher(dev)
    pending = {"resident": None, "e2e": None}
    loss_host = torch.empty(max(args.steps, args.warmup, 2) + 1, dtype=torch.float32).pin_memory()

    def resident_batch(i):
        s_dev, tok = resident[i % n_variants]
        return lambda: HostBatch(s_dev.clone(), tok)               # the builder writes fake activations in place

    def step_resident(i):
        if pending["resident"] is None:
            pending["resident"] = prefetch.submit(resident_batch(i))
        graph = prefetch.take(pending["resident"])
        pending["resident"] = prefetch.submit(resident_batch(i + 1))
        return step_fn(graph)

    def step_e2e(i):
        if pending["e2e"] is None:
            pending["e2e"] = prefetch.submit(hosts[i % n_variants])
        graph = prefetch.take(pending["e2e"])
        pending["e2e"] = prefetch.submit(hosts[(i + 1) % n_variants])   # pinned host -> device inside the timed region
        loss, _ = step_fn(graph)
        loss_host[i % loss_host.numel()].copy_(loss, non_blocking=True)  # D2H read of the step's result (read after the region)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.nvtx.range_push("timed")      # ncu --nvtx --nvtx-include "timed/" captures exactly this region
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.nvtx.range_pop()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for i in range(args.warmup):
        step_resident(i)
    for i in range(2):
        step_e2e(i)
    barrier()

    # ---- timed region 1: inputs resident in HBM; per-kernel CUDA events recorded live
    sampler = ClockSampler(local_rank) if rank == 0 else None
    _ffi.profiler.reset()
    _ffi.profiler.reserve(500 * args.steps)          # ~375 library calls per step: no event creation in the timed region
    _ffi.profiler.enabled = os.environ.get("PB200_NO_EVENTS") != "1"
    launches0 = _ffi.launch_counter["n"]
    if args.profiler_range:
        torch.cuda.profiler.start()
    total_ms = timed(step_resident, args.steps)
    if args.profiler_range:
        torch.cuda.profiler.stop()
    launches = _ffi.launch_counter["n"] - launches0
    _ffi.profiler.enabled = False
    summary = _ffi.profiler.summary()
    # ---- timed region 2: end to end from pinned host memory
    e2e_ms = total_ms if args.skip_e2e else timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None

    step_fn.close()
    secondary = None
    if not args.no_secondary:
        del step_fn, model, resident, pending
        secondary = secondary_measurements(args, dev, world, rank, pk)

    n_nodes = sum(h.tokens.size(0) for h in hosts) / n_variants
    graph0 = device_batch(hosts[0], dev)
    n_edges = graph0.num_edges
    global_batch = args.batch * world
    value = global_batch * args.steps / (total_ms * 1e-3)
    e2e_value = global_batch * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        st = graph0.structured if pb.ops.structured_enabled() else None
        rows = kernel_table(summary, int(n_nodes), int(n_edges), MODEL_CFG["d"], args.steps, args.precision, pk,
                            args.gcl_dropout, slots=3 if st is not None else 6,
                            n_rows=st.n_padded if st is not None else int(n_nodes),
                            act_bytes=2 if (args.precision == "bf16" and st is not None
                                            and pb.ops.bf16_activations_enabled()) else 4)
        ours_ms = sum(v["ms"] for v in summary.values()) / args.steps
        # dominant device kernel of the timed region: the three relational GEMM calls are one kernel
        # (pb::gemm_tcgen05_kernel); every other call family is its own kernel (pb_agg_bwd: dx + distance reduce)
        groups = {}
        for r in rows:
            if r["frac"] is None:
                continue
            key = "pb::gemm_tcgen05_kernel" if r["kernel"].startswith("pb_rgcn_gemm") else r["kernel"]
            groups.setdefault(key, []).append(r)
        top = None
        if groups:
            key, members = max(groups.items(), key=lambda kv: sum(m["ms_per_step"] for m in kv[1]))
            t_ms = sum(m["ms_per_step"] for m in members)
            work = sum(m["algorithmic_per_launch"] * m["launches_per_step"] for m in members)
            unit_scale = 1e12 if members[0]["bound"] == "tensor" else 1e9
            achieved = work / (t_ms * 1e-3) / unit_scale
            top = {"kernel": key, "bound": members[0]["bound"], "achieved": achieved, "peak": members[0]["peak"],
                   "unit": members[0]["unit"], "frac": achieved / members[0]["peak"], "ms_per_step": t_ms,
                   "algorithmic_per_launch": work / sum(m["launches_per_step"] for m in members),
                   "calls": [m["kernel"] for m in members]}
        hbm_rows = [r for r in rows if r["bound"] == "hbm"]
        hbm_bytes = sum(r["algorithmic_per_launch"] * 16 for r in hbm_rows)          # 16 GCL layers per step
        hbm_ms = sum(r["ms_per_step"] for r in hbm_rows)
        roofline = None
        if top:
            traffic = ncu_traffic().get(top["kernel"], {})
            roofline = {"kernel": top["kernel"], "abi_calls": top["calls"], "bound": top["bound"],
                        "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                        "traffic": traffic.get("bytes_per_launch"), "traffic_source": traffic.get("source"),
                        "peak_source": pk["source"], "algorithmic_per_launch": top["algorithmic_per_launch"],
                        "share_of_step": top["ms_per_step"] / (total_ms / args.steps),
                        "note": "largest share of the step among the path's device kernels; the HBM-bound kernels are "
                                "in `kernels` and summed in `mp_layer_hbm`; the largest of them is `largest_hbm_kernel`"}
            hb = max(hbm_rows, key=lambda r: r["ms_per_step"]) if hbm_rows else None
            if hb is not None:
                tr = ncu_traffic().get(hb["kernel"], {})
                roofline["largest_hbm_kernel"] = {
                    "kernel": hb["kernel"], "bound": "hbm", "achieved": hb["achieved"], "peak": hb["peak"], "unit": "GB/s",
                    "frac": hb["frac"], "traffic": tr.get("bytes_per_launch"), "traffic_source": tr.get("source"),
                    "algorithmic_per_launch": hb["algorithmic_per_launch"],
                    "share_of_step": hb["ms_per_step"] / (total_ms / args.steps)}
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only (the other ranks would wait)
            cpu = cpu_baseline_block(args)
        line = {
            "metric": "LMD16 train seqs/s", "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32 (tf32x3 tensor-core split)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": args.batch, "global_batch": global_batch, "n_bars": 16, "d": 512,
                       "gnn_n_layers": 8, "nodes_per_gpu": int(n_nodes), "edges_per_gpu": int(n_edges),
                       "gcl_dropout": args.gcl_dropout, "parallelism": f"dp{world}",
                       "operand_layout": "structured 4d (track-relation-sorted)" if st is not None else "generic 7d",
                       "activation_storage": "bf16 between the kernels of a GCN stack, fp32 arithmetic"
                       if (args.precision == "bf16" and st is not None and pb.ops.bf16_activations_enabled()) else "fp32",
                       "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "step": "device graph build + fwd + loss + bwd + NCCL grad all-reduce + Adam"},
            "e2e": {"value": e2e_value, "unit": "seq/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": hosts[0].nbytes, "d2h_bytes_per_step": 4 + 64,
                    "input": "pinned host buffers in the dataset's on-disk sample layout (int16 c_tensor [B,4,T,16,2], bool "
                             "s_tensor [B,4,T]); decode + graph build on the device (pb.decode_samples)"},
            "gpu_launches": launches, "roofline": roofline,
            "mp_layer_hbm": {"achieved": hbm_bytes / (hbm_ms * 1e-3) / 1e9 if hbm_ms else None, "peak": pk["hbm"],
                             "unit": "GB/s", "frac": hbm_bytes / (hbm_ms * 1e-3) / 1e9 / pk["hbm"] if hbm_ms else None,
                             "note": "all HBM-bound message-passing kernels (aggregate fwd/bwd, BN/ReLU/residual fwd/bwd)"},
            "kernels": rows, "our_kernels_ms_per_step": ours_ms, "cpu_baseline": cpu, "clocks": clocks,
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
