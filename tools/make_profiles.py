"""Turn gpurun_out/ captures into the tracked summaries under profiles/.

usage: python tools/make_profiles.py <tag> <bench.json> <launches.csv> <prof.ncu-rep> [<ref.json>] [name=<other.ncu-rep> ...]
(every name=<rep> adds profiles/<tag>_ncu_full_<name>.csv with the same headline metrics)
"""
import collections, csv, json, os, subprocess, sys
csv.field_size_limit(10**9)
tag, bench, launches, rep = sys.argv[1:5]
rest = sys.argv[5:]
extra_reps = [a.split("=", 1) for a in rest if "=" in a]
rest = [a for a in rest if "=" not in a]
ref = rest[0] if rest else None
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
os.makedirs(out, exist_ok=True)

# 1. bench line(s)
b = json.load(open(bench))
json.dump(b, open(os.path.join(out, "%s_bench.json" % tag), "w"), indent=1)
if ref:
    json.dump(json.load(open(ref)), open(os.path.join(out, "%s_bench_reference.json" % tag), "w"), indent=1)

# 2. launch list (ncu --metrics gpu__time_duration.sum), aggregated per kernel + raw list
rows = list(csv.reader(open(launches)))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, data = r, rows[i + 1:]
        break
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
raw = []
for r in data:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("dsrg::", "")
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1000.0 if u == "ns" else (v * 1000.0 if u == "ms" else v)
    raw.append((r[ix["ID"]], name, v))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, "%s_launches.csv" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one bench step (cold-cache, serialised: compare SHARES)\n")
    f.write("id,kernel,us\n")
    for r in raw:
        f.write("%s,%s,%.2f\n" % r)
lines = ["| kernel | launches | total µs | avg µs | share (ncu) | share (bench events) |", "|---|---|---|---|---|---|"]
tagmap = {"k_mf_tile<": "mf_tile", "k_mf_tile_hy": "mf_tile_hybrid", "k_mf_blur": None, "k_lattice_insert": "lattice_insert", "k_srg_label": "srg_label",
          "k_srg_emit": "srg_emit", "k_srg_merge": "srg_merge", "k_srg_flag": "srg_flag", "k_mf_zero": "mf_zero"}
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ev = ""
    for pre, tg in tagmap.items():
        if k.startswith(pre) and tg and tg in b.get("kernels", {}):
            ev = "%.1f %%" % (100 * b["kernels"][tg]["share"])
    lines.append("| `%s` | %d | %.1f | %.1f | %.1f %% | %s |" % (k, n, t, t / n, 100 * t / tot, ev))

# 3. ncu --set full headline metrics per captured kernel
def ncu_headlines(rep, dest):
    rawcsv = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(rawcsv.splitlines()))
    h, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]
    stalls = [x for x in h if x.startswith("smsp__average_warps_issue_stalled_") and x.endswith("_per_issue_active.ratio")]
    seen = {}
    with open(dest, "w") as f:
        f.write("kernel," + ",".join(want) + ",top_stalls\n")
        for r in rr[2:]:
            name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("dsrg::", "")
            key = name
            if seen.get(key, 0) >= 2:
                continue
            seen[key] = seen.get(key, 0) + 1
            vals = [r[h.index(w)] + " " + units[h.index(w)] if w in h else "" for w in want]
            st = sorted(((float(r[h.index(x)]), x.split("stalled_")[1].split("_per_issue")[0]) for x in stalls), reverse=True)[:5]
            f.write('"%s",' % name + ",".join('"%s"' % v for v in vals) + ',"' + " ".join("%s=%.2f" % (n, v) for v, n in st) + '"\n')


ncu_headlines(rep, os.path.join(out, "%s_ncu_full.csv" % tag))
for name, path in extra_reps:
    ncu_headlines(path, os.path.join(out, "%s_ncu_full_%s.csv" % (tag, name)))

with open(os.path.join(out, "%s_summary.md" % tag), "w") as f:
    f.write("# profile %s\n\n" % tag)
    f.write("bench: `%s` -> value %.1f %s, %.3f ms/step, e2e %.1f, launches/step %.0f, clocks %s\n\n" % (
        " ".join(["python bench.py", "--steps", str(b["steps"]), "--warmup", str(b["warmup"])]), b["value"], b["unit"], b["ms_per_step"],
        b["e2e"]["value"] if b.get("e2e") else float("nan"), b["gpu_launches"] / b["steps"], json.dumps(b.get("clocks"))))
    r = b["roofline"]
    f.write("roofline (dominant kernel `%s`): %.1f GB/s algorithmic = %.3f of %s peak %.1f GB/s; whole step %.1f GB/s = %.3f\n\n" % (
        r["kernel"], r["achieved"], r["frac"], r["peak_source"], r["peak"], r["step_algorithmic_GBs"], r["step_frac"]))
    if b.get("cpu_baseline"):
        f.write("cpu_baseline: %s\n\n" % json.dumps(b["cpu_baseline"]))
    f.write("## launch list of one step (ncu, serialised) vs bench CUDA-event shares\n\n" + "\n".join(lines) + "\n\n")
    f.write("## ncu --set full, headline metrics\n\nsee `%s_ncu_full.csv` (first two captures of each kernel)\n" % tag)
print("wrote profiles/%s_*" % tag)
