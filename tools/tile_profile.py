#!/usr/bin/env python3
"""Developer tool (GPU box): per-tile timing of the ray-tracing kernel for a bench workload."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import renderer_b200 as rb  # noqa: E402
from oracle import pyport  # noqa: E402  (model path staging only)
from bench import WORKLOADS  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
path = pyport.model_path(wl["model"])
s = rb.Scene(path).UpdateBoundingVolumeHierarchy(path + ".bvh")
g = rb.Renderer(0)
g.upload(s)
cam = rb.Orbit.cameras([10])[10]
f = rb.make_frame(wl["mode"], wl["W"], wl["H"], cam, flags=wl["flags"], ao_samples=wl["ao"] or 32, frame_index=10)
for _ in range(3):
    g.render(f)
t = g.tile_profile(f).astype(np.int64)
ok = (t[:, 0] > 0) & (t[:, 1] > 0)
st, en = t[ok, 0], t[ok, 1]
t0 = st.min()
dur = (en - st) / 1e3
span = (en.max() - t0) / 1e3
order = np.argsort(dur)[::-1]
res = {"tiles": int(ok.sum()), "kernel_span_us": float(span), "sum_tile_us": float(dur.sum()),
       "max_tile_us": float(dur.max()), "p99_us": float(np.percentile(dur, 99)), "p90_us": float(np.percentile(dur, 90)),
       "median_us": float(np.median(dur)),
       "tiles_over_50us": int((dur > 50).sum()), "tiles_over_20us": int((dur > 20).sum()),
       "start_of_longest_us": [float((st[i] - t0) / 1e3) for i in order[:5]],
       "dur_of_longest_us": [float(dur[i]) for i in order[:5]]}
# concurrency timeline: warps busy (tiles in flight) sampled every 5% of the span
ts = np.linspace(0, span, 21)
res["tiles_in_flight"] = [int((((st - t0) / 1e3 <= x) & ((en - t0) / 1e3 > x)).sum()) for x in ts]
res["last_start_us"] = float((st.max() - t0) / 1e3)
print(json.dumps(res))
