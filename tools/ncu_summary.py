"""Summarise an .ncu-rep: headline metrics (raw page) + SASS-region profile (source page)."""
import csv, subprocess, sys
csv.field_size_limit(10**9)
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w); print("%-75s %s %s" % (w, r[i], units[i]))
    st = sorted(((float(r[hdr.index(h)]), h) for h in stalls), reverse=True)[:7]
    print("top stalls (warps per issue):", ", ".join("%s=%.2f" % (h.split('stalled_')[1].split('_per_issue')[0], v) for v, h in st))
if len(sys.argv) > 2 and sys.argv[2] == "--sass":
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    seg = rows[starts[0] + 1:starts[1]]
    hdr, data = seg[0], seg[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    f = lambda r, k: float(r[ix[k]]) if r[ix[k]] not in ('', None) else 0.0
    ti = sum(f(r, 'Instructions Executed') for r in data); ts = sum(f(r, '# Samples') for r in data)
    print("SASS regions (40 instr each): n=%d inst=%.0f samples=%.0f" % (len(data), ti, ts))
    for c in range(0, len(data), 40):
        d = data[c:c + 40]
        ii = sum(f(r, 'Instructions Executed') for r in d); ss = sum(f(r, '# Samples') for r in d)
        ops = {}
        for r in d:
            t = r[1].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] = ops.get(op, 0) + 1
        if ii / ti > 0.004 or ss / ts > 0.004:
            print("%5d-%5d inst %5.1f%% samp %5.1f%% | %s" % (c, c + 40, 100 * ii / ti, 100 * ss / ts,
                  ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:7])))
