#!/usr/bin/env python3
"""Developer tool (GPU box): per-warp timeline of rt_primary_kernel (B200R_WARP_PROFILE=1)."""
import json, os, sys
import numpy as np
os.environ["B200R_WARP_PROFILE"] = "1"
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
w = out.reshape(-1, 4).astype(np.int64)
w = w[w[:, 0] > 0]
t0 = w[:, 0].min()
beg, end, rays, shadows = (w[:, 0] - t0) / 1e3, (w[:, 1] - t0) / 1e3, w[:, 2] & 0xfffff, (w[:, 2] >> 20) & 0xfffff
rounds, refills, rounds_after, drained_us = w[:, 3] & 0xffff, (w[:, 3] >> 16) & 0xfff, (w[:, 3] >> 28) & 0xfff, (w[:, 3] >> 40) / 10.0
dur = end - beg
res = {"warps": int(len(w)), "span_us": float(end.max()), "begin_max_us": float(beg.max()),
       "end_pct_us": {p: float(np.percentile(end, p)) for p in (10, 50, 90, 99, 100)},
       "rays_per_warp": {"min": int(rays.min()), "median": float(np.median(rays)), "max": int(rays.max()), "sum": int(rays.sum())},
       "rounds_per_warp": {"median": float(np.median(rounds)), "max": int(rounds.max())},
       "refills_per_warp": {"median": float(np.median(refills)), "max": int(refills.max())},
       "us_per_round_median": float(np.median(dur / np.maximum(rounds, 1))),
       "drained_us": {p: float(np.percentile(drained_us, p)) for p in (1, 50, 99)},
       "rounds_after_drain": {"median": float(np.median(rounds_after)), "p90": float(np.percentile(rounds_after, 90)), "max": int(rounds_after.max())},
       "shadow_rays_per_warp": {"median": float(np.median(shadows)), "max": int(shadows.max()), "sum": int(shadows.sum())},
       "donated_subtrees": int((w[:, 2] >> 40).sum())}
print(json.dumps(res))
