#!/usr/bin/env python3
"""Developer tool (GPU box): per-phase iteration / lane statistics of rt_pool_kernel for one frame (switch pool_stats).
    python tools/pool_stats.py [c2|c3|c5] [frame] [row_step]      row_step P: only rows 0, P, 2P, .. (what rank 0 of P renders)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import renderer_b200 as rb
from oracle import pyport
from bench import WORKLOADS
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]; k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model = pyport.model_path(wl["model"])
scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
g = rb.Renderer(0); g.upload(scene)
P = int(sys.argv[3]) if len(sys.argv) > 3 else 1
f = rb.make_frame(wl["mode"], wl["W"], wl["H"], rb.Orbit.cameras([k])[k], flags=wl["flags"], ao_samples=wl["ao"] or 32, frame_index=k,
                  row_first=0, row_step=P)
g.render(f)
import ctypes as C
L = rb.lib()
g.set_switch("pool_stats", 1)
g.render(f)
c = list(g.counters().values())
names = ["inner", "leaf", "resolve", "refill", "guard"]
out = {n: {"iterations": c[2 * i], "lanes": c[2 * i + 1], "lanes_per_iteration": round(c[2 * i + 1] / max(c[2 * i], 1), 2)} for i, n in enumerate(names)}
out["inner_pops_dropped"] = c[10]; out["cold_pops"] = c[8] >> 32; out["guard"]["iterations"] &= 0xffffffff
out["max_iterations_of_a_warp"] = c[9]; out["guard"]["lanes"] = 0; out["warps"] = 148 * 3 * 8; out["kernel_ms_with_stats"] = g.last_kernel_ms()[0]
g.set_switch("pool_stats", 0)
ts = []
for _ in range(5):
    g.render(f); ts.append(g.last_kernel_ms()[0])
out["kernel_ms"] = min(ts); out["row_step"] = P
print(json.dumps(out))
