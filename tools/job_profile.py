#!/usr/bin/env python3
"""Developer tool (GPU box): job-length statistics of rt_wave_kernel (PROF build; B200R_WARP_PROFILE + B200R_JOB_PROFILE).
Prints per-phase lane utilisation, the histogram of steps (inner + triangle iterations) per job kind, and the long jobs."""
import json, os, sys
import numpy as np
os.environ["B200R_WARP_PROFILE"] = "1"
os.environ["B200R_JOB_PROFILE"] = "1"
sys.path.insert(0, ".")
import renderer_b200 as rb
from oracle import pyport
from bench import WORKLOADS
import ctypes as C
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
path = pyport.model_path(wl["model"])
s = rb.Scene(path).UpdateBoundingVolumeHierarchy(path + ".bvh")
g = rb.Renderer(0); g.upload(s)
cam = rb.Orbit.cameras([10])[10]
f = rb.make_frame(wl["mode"], wl["W"], wl["H"], cam, flags=wl["flags"], ao_samples=wl["ao"] or 32, frame_index=10)
for _ in range(3): g.render(f)
n = C.c_uint32(); L = rb.lib()
L.b200r_get_tile_profile(g._ctx, None, 0, C.byref(n))
out = np.zeros((n.value, 2), dtype=np.uint64)
L.b200r_get_tile_profile(g._ctx, out.ctypes.data, n.value, C.byref(n))
u = out.reshape(-1)
B = 131072
w = u[:B].reshape(-1, 4).astype(np.int64); w = w[w[:, 0] > 0]
t0 = int(w[:, 0].min()); span = (int(w[:, 1].max()) - t0) / 1e3
iters = u[B:B + 4].astype(np.int64); lanes = u[B + 4:B + 8].astype(np.int64)
names = ["inner", "leaf", "fin", "shade"]
res = {"span_us": span, "phases": {nm: {"iters": int(i), "lanes_per_iter": float(l) / max(int(i), 1)} for nm, i, l in zip(names, iters, lanes)}}
hist = u[B + 16:B + 16 + 8 * 64].astype(np.int64).reshape(8, 64)
kinds = ["prim-miss", "prim-hit", "shadow-lit", "shadow-blocked", "don-prim-miss", "don-prim-hit", "don-shadow-lit", "don-shadow-blocked"]
hs = {}
for k, nm in enumerate(kinds):
    h = hist[k]; tot = int(h.sum())
    if not tot: continue
    steps = (np.arange(64) * 8 + 4)
    cum = np.cumsum(h) / tot
    hs[nm] = {"jobs": tot, "mean_steps": float((h * steps).sum() / tot), "p50": int(steps[np.searchsorted(cum, .5)]), "p90": int(steps[np.searchsorted(cum, .9)]),
              "p99": int(steps[np.searchsorted(cum, .99)]), "max_bin": int(steps[np.nonzero(h)[0].max()]), "total_steps": int((h * steps).sum())}
res["jobs"] = hs
nlog = int(u[B + 1024]); res["long_jobs"] = nlog
rec = u[B + 1026:B + 1026 + 4 * min(nlog, 30000)].astype(np.int64).reshape(-1, 4)
if len(rec):
    pix = rec[:, 0] & 0xffffffff; steps = rec[:, 0] >> 32; kind = rec[:, 1] & 0xff; tstart = (rec[:, 1] >> 32) / 1e3
    tend = rec[:, 2] / 1e3; wbeg = (rec[:, 3] - t0) / 1e3
    x = pix & 0xffff; y = pix >> 16
    order = np.argsort(-(tend + wbeg))[:25]
    res["latest_long_jobs"] = [{"x": int(x[i]), "y": int(y[i]), "steps": int(steps[i]), "kind": kinds[int(kind[i])], "start_us": round(float(tstart[i] + wbeg[i]), 1),
                                "end_us": round(float(tend[i] + wbeg[i]), 1), "us_per_step": round(float((tend[i] - tstart[i]) / max(int(steps[i]), 1)), 2)} for i in order]
    res["long_by_kind"] = {kinds[k]: int((kind == k).sum()) for k in range(8) if (kind == k).any()}
    res["long_rows"] = {"y_min": int(y.min()), "y_max": int(y.max()), "y_hist_16": np.histogram(y, bins=16, range=(400, 700))[0].tolist()}
    res["long_us_per_step"] = {"median": float(np.median((tend - tstart) / np.maximum(steps, 1))), "p10": float(np.percentile((tend - tstart) / np.maximum(steps, 1), 10))}
print(json.dumps(res))
