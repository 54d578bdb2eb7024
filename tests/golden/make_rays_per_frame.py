#!/usr/bin/env python3
"""Rays (BVH_IntersectTriangles invocations: primary + shadow + reflection + AO) per orbit frame of every ray-traced bench
workload, counted by the CPU restatement (oracle/port, whose counters equal the instrumented reference's).
Written to tests/golden/rays_per_frame.json so that `bench.py --impl reference` can turn the reference's fps into Mrays/s
without loading anything of the product.      python tests/golden/make_rays_per_frame.py [workload ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import renderer_b200 as rb
from oracle import pyport
from bench import WORKLOADS
OUT = os.path.join(ROOT, "tests", "golden", "rays_per_frame.json")
STRIDE = {"c2": 1, "c2r": 4, "c3": 8, "c5": 16}
N = 128
data = json.load(open(OUT)) if os.path.exists(OUT) else {}
for name in (sys.argv[1:] or list(STRIDE)):
    wl = WORKLOADS[name]
    model = pyport.model_path(wl["model"])
    scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
    cams = rb.Orbit.cameras(range(N))
    d = {}
    for k in range(0, N, STRIDE[name]):
        f = rb.make_frame(wl["mode"], wl["W"], wl["H"], cams[k], flags=wl["flags"], ao_samples=wl["ao"] or 32, frame_index=k)
        _, c = pyport.render(scene, f, counters=True)
        d[str(k)] = c["rays_primary"] + c["rays_shadow"] + c["rays_reflection"] + c["rays_ao"]
        print(name, k, d[str(k)], flush=True)
    data[name] = d
    json.dump(data, open(OUT, "w"), indent=0, sort_keys=True)
