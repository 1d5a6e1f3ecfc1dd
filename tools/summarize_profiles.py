"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run in the container).

    python tools/summarize_profiles.py r01
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
if len(sys.argv) > 2:            # on the GPU box: summaries next to the captures (only gpurun_out/ travels back)
    dst = os.path.join(ROOT, sys.argv[2])
    os.makedirs(dst, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launch_shares():
    path = os.path.join(src, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[start], rows[start + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        name = r[ki].split("(")[0][:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "pb::" in k)
    out = [f"# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none  (bench.py --steps 1, LMD16, "
           f"per-GPU batch 256, bf16): every launch of ONE timed step",
           f"# {len(data)} launches, {tot / 1e3:.2f} ms serialised cold-cache device time; kernels of libpolyphemus_b200 (pb::*): "
           f"{ours / 1e3:.2f} ms = {ours / tot:.1%}. Compare SHARES with bench.py's live CUDA-event numbers, not absolutes.",
           "kernel,launches,total_us,share"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:60]:
        out.append(f"\"{k}\",{n},{t:.1f},{t / tot:.4f}")
    open(os.path.join(dst, f"{tag}_launch_shares.csv"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:22]))


def rep_summaries():
    for rep in sorted(glob.glob(os.path.join(src, f"prof_*_{tag}.ncu-rep"))):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        lines = [f"# ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}); one row block per captured launch"]
        for r in rows[2:]:
            lines.append(f"kernel: {r[idx['Kernel Name']][:110]}")
            for k in KEYS:
                if k in idx and r[idx[k]] not in ("", "n/a"):
                    lines.append(f"  {k:88s} {r[idx[k]][:24]:>24s} {units[idx[k]]}")
        name = os.path.basename(rep).replace(".ncu-rep", ".txt").replace("prof_", f"{tag}_ncu_").replace(f"_{tag}.txt", ".txt")
        open(os.path.join(dst, name), "w").write("\n".join(lines) + "\n")
        print("wrote", name)


def traffic_json():
    """profiles/traffic.json: measured DRAM bytes per launch of each kernel family (bench.py's roofline.traffic)."""
    import json
    import re

    families = {   # ABI call -> (summary file, kernel-name prefixes whose launches make up one call)
        "pb_agg_fwd": ("aggfwd", ["agg_fwd_pipe_kernel", "agg_fwd_kernel"]),
        "pb_agg_bwd": ("aggbwd", ["agg_bwd_dx_kernel"]),
        "pb_agg_bwd_fused": ("aggbwdring", ["agg_bwd_ring_kernel"]),
        "pb_dist_reduce": ("distred", ["dist_reduce_kernel"]),
        "pb_rgcn_gemm_fwd": ("gemm", ["gemm_tcgen05_kernel"]),
        "pb::gemm_tcgen05_kernel": ("gemm", ["gemm_tcgen05_kernel"]),     # mean of the captured launches (fwd, bwd_weight, bwd_data)
        "pb_bn_relu_res_fwd": ("bn", ["bn_apply_kernel"]),
        "pb_bn_stats": ("bn", ["bn_stats_partial_kernel", "bn_stats_finalize_kernel"]),
        "pb_bn_relu_res_bwd": ("bn", ["bn_bwd_partial_kernel", "bn_bwd_apply_kernel"]),
        "pb_dropout_bits": ("dropbits", ["dropout_bits_kernel"]),
        "pb_ce_rows": ("ce", ["ce_rows_kernel"]),
        "pb_chord_embed_fwd": ("chord", ["chord_embed_fwd_kernel"]),
        "pb_bar_pool": ("pool", ["bar_pool_fwd_kernel", "bar_pool_bwd_kernel"]),
        "pb_rows_scatter": ("rows", ["rows_permute_kernel"]),
    }
    out = {}
    for call, (stem, prefixes) in families.items():
        path = os.path.join(dst, f"{tag}_ncu_{stem}.txt")
        if not os.path.exists(path):
            continue
        per_kernel, cur = {}, None
        for line in open(path):
            if line.startswith("kernel:"):
                cur = line.split("kernel:")[1].strip()
                cur = re.sub(r"^void ", "", cur).split("<")[0].split("(")[0].split("::")[-1]
                per_kernel.setdefault(cur, []).append(0.0)
            elif cur and ("dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line):
                val, unit = line.split()[-2:]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                per_kernel[cur][-1] += float(val) * mult
        total, found = 0.0, []
        for pre in prefixes:
            if pre in per_kernel:
                total += sum(per_kernel[pre]) / len(per_kernel[pre])
                found.append(pre)
        if found:
            out[call] = {"bytes_per_launch": total, "kernels": found,
                         "source": f"profiles/{tag}_ncu_{stem}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    if "pb_agg_bwd" in out and "pb_dist_reduce" in out:     # the deterministic path = both kernels
        out["pb_agg_bwd"]["bytes_per_launch"] += out["pb_dist_reduce"]["bytes_per_launch"]
        out["pb_agg_bwd"]["kernels"] += out["pb_dist_reduce"]["kernels"]
        out["pb_agg_bwd"]["source"] += f" + profiles/{tag}_ncu_distred.txt"
    out.pop("pb_dist_reduce", None)
    path = os.path.join(dst, "traffic.json")
    merged = json.load(open(path)) if os.path.exists(path) else {}      # entries of earlier rounds stay unless re-measured
    merged.update(out)
    out = merged
    json.dump(out, open(path, "w"), indent=1)
    print("wrote traffic.json:", {k: round(v["bytes_per_launch"] / 1e6, 1) for k, v in out.items()}, "MB")


launch_shares()
rep_summaries()
traffic_json()
