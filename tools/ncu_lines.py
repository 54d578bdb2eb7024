#!/usr/bin/env python3
"""Per-source-line hot spots of an ncu capture (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [topN] [kernel-id]
Prints lines sorted by warp instructions executed, with average active threads and stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; recs = []; kernel_no = -1
want_kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0] == "": continue          # a SASS row under the source line
    d = {}
    for h, v in zip(hdr, r):
        d.setdefault(h, v)
    try:
        ie = int(d.get("Instructions Executed", "0") or 0); te = int(d.get("Thread Instructions Executed", "0") or 0)
        sm = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    recs.append((fname, int(d["Line No"]), ie, te, sm, d.get("Source", "")[:110]))
# merge duplicates (same file/line appears once per kernel result); keep totals
agg = {}
for f, ln, ie, te, sm, s in recs:
    k = (f, ln)
    a = agg.setdefault(k, [0, 0, 0, s]); a[0] += ie; a[1] += te; a[2] += sm
tot_ie = sum(a[0] for a in agg.values()); tot_te = sum(a[1] for a in agg.values()); tot_sm = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot_ie}  thread-inst {tot_te}  avg threads/inst {tot_te/max(tot_ie,1):.2f}  samples {tot_sm}")
print(f"{'file:line':28s} {'inst%':>6s} {'thr/inst':>8s} {'smp%':>6s}  source")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f+':'+str(ln):28s} {100*a[0]/tot_ie:6.2f} {a[1]/max(a[0],1):8.2f} {100*a[2]/max(tot_sm,1):6.2f}  {a[3].strip()}")
