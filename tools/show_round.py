"""Print the bench line and the kernel table of the last GPU round (gpurun_out/t_bench.log, t_prof.log)."""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for line in open(os.path.join(root, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "t_bench.log")):
    if line.startswith("{"):
        j = json.loads(line)
        print("value", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 2), "e2e", round(j["e2e"]["value"], 1) if j.get("e2e") else None,
              "launches", j["gpu_launches"], "ours ms/step", round(j.get("our_kernels_ms_per_step", 0), 2), "clocks", j.get("clocks"))
        for k in j["kernels"]:
            print(f"  {k['kernel']:32s} {k['ms_per_step']:7.2f} ms/step  avg {k['avg_ms']:.3f} ms  {(k['achieved'] or 0):8.1f} {k['unit']}  frac {(k['frac'] or 0):.2f}  x{k['launches_per_step']}")
