#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ evidence run (scripts/gpu_final.sh) into the tracked summaries under profiles/:
  profiles/<prefix>_bench.json, _bench_reference.json   the two bench lines
  profiles/<prefix>_ncu_launches_summary.txt            per-kernel share of one step (ncu launch list)
  profiles/<prefix>_ncu_launches.csv                    the raw launch list
  profiles/<prefix>_ncu_full_summary.txt                key `--set full` metrics per kernel
  profiles/traffic.json                                 dram read+write bytes per launch (bench.py's roofline.traffic)
  profiles/limiters.json                                issue-slot / DRAM / occupancy figures per stage (bench.py's roofline.ncu)
usage: scripts/summarize_profiles.py gpurun_out/<tag> <prefix>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

src, prefix = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

for name in ("bench.json", "bench_reference.json", "sass_excerpt.txt", "cub_sort_yardstick.txt"):
    f = os.path.join(src, name)
    if os.path.exists(f) and os.path.getsize(f):
        shutil.copy(f, os.path.join(P, f"{prefix}_{name}"))

# ---- launch list
lf = os.path.join(src, "launches.csv")
if os.path.exists(lf):
    shutil.copy(lf, os.path.join(P, f"{prefix}_ncu_launches.csv"))
    rows = list(csv.DictReader([l for l in open(lf) if not l.startswith("==")]))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "sample_reduce_fwd" in n]
    s, e = idx[-2], idx[-1]
    agg = collections.OrderedDict()
    tot = ours = 0.0
    for r in rows[s:e]:
        n = r["Kernel Name"]
        v = float(r["Metric Value"]) / 1000.0
        key = n.split("(")[0][:80]
        agg.setdefault(key, [0.0, 0])
        agg[key][0] += v
        agg[key][1] += 1
        tot += v
        if "cg::" in n or "ssimk::" in n:
            ours += v
    with open(os.path.join(P, f"{prefix}_ncu_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 700 python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline\n")
        f.write(f"one step: {e - s} launches, {tot:.1f} us serialised (cold-cache, profiler-serialised times; shares are what matter)\n")
        f.write(f"libcurvegs kernels: {ours:.1f} us ({100 * ours / tot:.1f}% of the step); torch glue kernels: {tot - ours:.1f} us\n\n")
        for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{v:9.1f} us {100 * v / tot:5.1f}%  x{c:3d}  {k}\n")

# ---- full capture
rep = os.path.join(src, "full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    units = dict(zip(hdr, rows[1]))
    traffic = {}
    limiters = {}     # per stage: what ncu says keeps the kernel busy (read by bench.py next to roofline.traffic)
    stage_of = {"blend_bwd": "blend_bwd", "blend_bwd_ring": "blend_bwd", "blend_fwd": "blend_fwd",
                "tile_ranges": "tile_ranges", "emit_keys": "emit_keys",
                "preprocess_fwd": "preprocess_fwd", "preprocess_bwd": "preprocess_bwd"}
    seen = collections.Counter()
    with open(os.path.join(P, f"{prefix}_ncu_full_summary.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on (one steady-state step of bench.py --views 2, C4)\n")
        f.write("per-launch values; times under the profiler are cold-cache/serialised\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            kn = d["Kernel Name"]
            base = kn.split("(")[0].split("<")[0].split("::")[-1].split()[-1].strip()
            seen[base] += 1
            f.write(f"{kn[:110]}   grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for w in want:
                if w in d:
                    f.write(f"    {w:75s} {d[w]:>18s} {units.get(w, '')}\n")
            try:
                mb = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(units.get("dram__bytes_read.sum", "Mbyte"), 1e6)
                if base in stage_of and stage_of[base] not in traffic:
                    traffic[stage_of[base]] = int(mb * scale)
                    limiters[stage_of[base]] = {
                        "issue_active_pct": round(float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), 1),
                        "dram_pct_of_peak": round(float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]), 1),
                        "warps_active_pct": round(float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]), 1),
                        "registers": int(float(d["launch__registers_per_thread"])),
                        "inst_executed": int(float(d["smsp__inst_executed.sum"])),
                        "source": f"profiles/{prefix}_ncu_full_summary.txt"}
                if base == "sort_onesweep_pass":
                    traffic["radix_sort"] = traffic.get("radix_sort", 0) + int(mb * scale)
            except Exception:
                pass
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    json.dump(limiters, open(os.path.join(P, "limiters.json"), "w"), indent=1)
    print("traffic", traffic)
print("wrote summaries for", src, "->", P)
