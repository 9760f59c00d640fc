"""Turn gpurun_out/ ncu dumps into the tracked summaries under profiles/.

    python tools/make_profiles.py r01        (expects gpurun_out/r01_launches.csv and
                                              gpurun_out/r01_agg_stream_raw.csv = `ncu -i ... --page raw --csv`)
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and shares -------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", tag + "_launches.csv"))) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, tag + "_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 600   python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    f.write("(cold-cache, serialised launches: compare SHARES, not absolutes)\n\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-100s n=%4d %12.1f us %6.2f%%\n" % (k[:100], a[0], a[1], 100 * a[1] / tot))
# the csv itself is small: keep it
with open(os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")) as f, open(os.path.join(out_dir, tag + "_launches.csv"), "w") as g:
    g.write(f.read())

# ---- full capture of the hot kernel ---------------------------------------------------------------
raw = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", tag + "_agg_stream_raw.csv"))))
h, units = raw[0], raw[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]
names = {2: "fwd_shared (layer 1, X shared by the 16 samples)", 3: "fwd (layer 2, per-sample X)", 4: "fwd (layer 3)"}
traffic = {}
with open(os.path.join(out_dir, tag + "_agg_stream_full.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:agg_stream -s 18 -c 3   python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    for k in range(2, len(raw)):
        r = raw[k]
        f.write("\n== launch %d: %s\n" % (k - 2, names.get(k, "")))
        vals = {}
        for w in want:
            for i, c in enumerate(h):
                if c == w:
                    f.write("%-72s %s %s\n" % (w, r[i], units[i]))
                    vals[w] = (r[i], units[i])
        for i, c in enumerate(h):
            if "issue_stalled" in c and c.endswith("per_issue_active.ratio") and "not_issued" not in c:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.15:
                    f.write("   stall %-50s %s\n" % (c.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), r[i]))

        def to_bytes(x):
            v, u = x
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        traffic[k] = to_bytes(vals["dram__bytes_read.sum"]) + to_bytes(vals["dram__bytes_write.sum"])
json.dump({"fwd_shared": traffic.get(2), "fwd": traffic.get(3), "bwd": traffic.get(3),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of stag::agg_stream_kernel, ncu --set full, "
                   "bench.py arxiv shape S=16; the transposed launch (bwd) runs the same kernel on the CSR"},
          open(os.path.join(out_dir, tag + "_traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, tag + "_launches_summary.txt")).read())
print(json.load(open(os.path.join(out_dir, tag + "_traffic.json"))))
