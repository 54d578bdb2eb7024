#!/usr/bin/env python3
"""Condense a GPU session's artefacts (gpurun_out/<tag>_*) into small tracked summaries under profiles/.

    python tools/summarize_profiles.py r01f [c2]
Writes profiles/<tag>_bench_<wl>.json, profiles/<tag>_launches_<wl>.csv (per-kernel table from the ncu
gpu__time_duration pass), profiles/<tag>_ncu_<wl>.txt (key metrics of the `--set full` capture of the dominant kernel,
incl. dram__bytes_read/write -> traffic) and copies the tile/warp profiles. Updates profiles/traffic.json.
"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

for name in (f"{tag}_bench_{wl}.json", f"{tag}_tiles_{wl}.json", f"{tag}_warps_{wl}.json", f"{tag}_pytest.log", f"{tag}_smoke.log",
             f"{tag}_smi.txt"):
    src = os.path.join(G, name)
    if os.path.exists(src) and os.path.getsize(src) > 0:
        shutil.copyfile(src, os.path.join(P, name))

lc = os.path.join(G, f"{tag}_launches_{wl}.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    d = defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) > vi:
            d[(r[ki].split("(")[0].replace("void ", "").replace("b200r::", ""), r[gi], r[bi])].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(P, f"{tag}_launches_{wl}.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --workload %s --steps 4 --warmup 3\n" % wl)
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write("kernel,grid,block,launches,avg_us,total_us,share\n")
        for (k, g, b), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",\"{g}\",\"{b}\",{len(v)},{sum(v)/len(v)/1e3:.2f},{sum(v)/1e3:.1f},{sum(v)/tot:.3f}\n")

rep = os.path.join(G, f"{tag}_prof_{wl}.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) > 2:
        hdr, units = rows[0], rows[1]
        want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
                "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct",
                "smsp__thread_inst_executed_per_inst_executed.ratio", "_per_issue_active.ratio", "sm__throughput.avg.pct",
                "sm__cycles_active.avg", "gpu__dram_throughput", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
        traffic = None
        with open(os.path.join(P, f"{tag}_ncu_{wl}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on  (capture: gpurun_out/{tag}_prof_{wl}.ncu-rep, not tracked)\n")
            for r in rows[2:]:
                rd, wr = None, None
                for h, u, v in zip(hdr, units, r):
                    if any(w in h for w in want) and "pcsamp" not in h and "TriageCompute" not in h and ".max." not in h and ".min." not in h \
                            and ".sum.p" not in h and ".per_second" not in h:
                        f.write(f"{h} [{u}] = {v}\n")
                    if h == "dram__bytes_read.sum":
                        rd = float(v.replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                    if h == "dram__bytes_write.sum":
                        wr = float(v.replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                if rd is not None and wr is not None:
                    traffic = rd + wr
                f.write("----\n")
        if traffic is not None:
            tp = os.path.join(P, "traffic.json")
            t = json.load(open(tp)) if os.path.exists(tp) else {}
            t[wl] = traffic
            t[wl + "_source"] = f"{tag}_ncu_{wl}.txt (dram__bytes_read.sum + dram__bytes_write.sum, last captured launch of the dominant kernel)"
            json.dump(t, open(tp, "w"), indent=1)
print("ok")
