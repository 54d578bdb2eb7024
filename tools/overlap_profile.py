#!/usr/bin/env python3
"""Developer tool (GPU box): how do ray-traced frames that are IN FLIGHT TOGETHER share the machine?

    python tools/overlap_profile.py [workload] [frames_in_flight] [row_step]

Renders 4*depth orbit frames through renderer_b200.dist.FramePipeline (world 1) with B200R_WARP_PROFILE=1 and prints, for the last
`depth` frames (one per scratch slot), when the warps of each frame's persistent kernel began and ended on ONE time axis
(globaltimer), plus a coarse timeline: warps alive per frame in 10 us buckets. row_step > 1 renders every row_step-th row only
(what one rank of row_step GPUs does). Answers DESIGN.md section 8 item 2: does frame i+1 start when frame i's job queue runs dry,
or only when its last warps retire?
"""
import json
import os
import sys

import numpy as np

os.environ["B200R_WARP_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import renderer_b200 as rb
from bench import WORKLOADS
from oracle import pyport          # model staging paths only
from renderer_b200.dist import FramePipeline

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
row_step = int(sys.argv[3]) if len(sys.argv) > 3 else 1
path = pyport.model_path(wl["model"])
scene = rb.Scene(path).UpdateBoundingVolumeHierarchy(path + ".bvh")
gpu = rb.Renderer(0)
gpu.upload(scene)
n = 4 * depth
cams = rb.Orbit.cameras(range(n))
frames = [rb.make_frame(wl["mode"], wl["W"], wl["H"], cams[k], flags=wl["flags"], ao_samples=wl["ao"] or 32, frame_index=k,
                        row_first=0, row_step=row_step) for k in range(n)]
pipe = FramePipeline(gpu, wl["W"], wl["H"], depth=depth)
for f in frames:
    pipe.submit(f)
pipe.drain()

L = rb.lib()
per_slot = []
for d in range(depth):
    cnt = C.c_uint32()
    L.b200r_get_warp_profile(gpu._ctx, d, None, 0, C.byref(cnt))
    rec = np.zeros((cnt.value, 4), dtype=np.uint64)
    if cnt.value:
        L.b200r_get_warp_profile(gpu._ctx, d, rec.ctypes.data, cnt.value, C.byref(cnt))
    rec = rec.astype(np.int64)
    per_slot.append(rec[rec[:, 0] > 0])
t0 = min(int(r[:, 0].min()) for r in per_slot if len(r))
order = sorted(range(depth), key=lambda d: int(per_slot[d][:, 0].min()) if len(per_slot[d]) else 0)
out = {"workload": wl["desc"], "frames_in_flight": depth, "row_step": row_step, "frames": []}
t_end = max(int(r[:, 1].max()) for r in per_slot if len(r)) - t0
buckets = np.arange(0, t_end / 1e3 + 10, 10.0)
for d in order:
    r = per_slot[d]
    beg, end = (r[:, 0] - t0) / 1e3, (r[:, 1] - t0) / 1e3
    drained = beg.min() + (r[:, 3] >> 40) / 10.0
    alive = [int(((beg <= b + 10) & (end >= b)).sum()) for b in buckets]
    out["frames"].append({"slot": d, "warps": int(len(r)),
                          "begin_us": {p: float(np.percentile(beg, p)) for p in (0, 10, 50, 90, 100)},
                          "end_us": {p: float(np.percentile(end, p)) for p in (0, 10, 50, 90, 99, 100)},
                          "queue_dry_us_median": float(np.median(drained)),
                          "warps_alive_per_10us": alive})
print(json.dumps(out))
