#!/usr/bin/env python3
"""Developer tool (GPU box): wall time of b200r_build_bvh vs the host builder, per model."""
import json, os, sys, time
sys.path.insert(0, ".")
import renderer_b200 as rb
from oracle import pyport
g = rb.Renderer(0)
out = {}
for m in (sys.argv[1:] or ["torus.ply", "trainColor.tri", "chessboard.tri", "dragon_vis.ply", "statue.ply"]):
    s = rb.Scene(pyport.model_path(m))
    s.bvh_bytes_from_device_build(g)                       # warm-up (allocations, module load)
    t = time.time(); b, d = s.bvh_bytes_from_device_build(g); dev = time.time() - t
    t = time.time(); s.UpdateBoundingVolumeHierarchy(None, forceRecalc=True); host = time.time() - t
    out[m] = {"tris": s.n_triangles, "depth": d, "device_ms": round(dev * 1e3, 1), "host_builder_ms": round(host * 1e3, 1), "same": b == s.bvh_bytes()}
print(json.dumps(out))
