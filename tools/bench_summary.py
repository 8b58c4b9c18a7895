import json, sys
d = [json.loads(l[l.index("{"):]) for l in open(sys.argv[1]) if "{" in l and "metric" in l][-1]
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"] if d.get("e2e") else None, d["clocks"])
r = d["roofline"]; print({k: r[k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "share_of_step", "step_frac")})
for k, v in d["kernels"].items():
    print("  %-20s %8.3f ms/step  %5.1f launches  %5.1f%%" % (k, v["ms_per_step"], v["launches_per_step"], 100 * v["share"]))
if d.get("cpu_baseline"): print(d["cpu_baseline"])
