"""Print the per-kernel table of a bench.py JSON line (stdin or file)."""
import json, sys
line = [l for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin) if l.startswith("{")][-1]
j = json.loads(line)
print(f"value {j['value']:.0f} seq/s  {j['ms_per_step']:.2f} ms/step   e2e {j['e2e']['value']:.0f} ({j['e2e'].get('ms_per_step', 0):.2f} ms)  launches {j['gpu_launches']}")
print(f"mp_layer_hbm frac {j['mp_layer_hbm']['frac']:.3f}   ours {j['our_kernels_ms_per_step']:.2f} ms/step")
for k in j["kernels"]:
    fr = "  -  " if k["frac"] is None else f"{k['frac']:.3f}"
    print(f"  {k['kernel']:34s} {k['avg_ms']*1e3:8.1f} us x{k['launches_per_step']:5.1f} = {k['ms_per_step']:6.2f} ms  frac {fr}")
