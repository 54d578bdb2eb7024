#!/usr/bin/env python3
"""Small frames through every device path, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize.py
C2-type frame (fused pool kernel), generic frame (AO + reflections: wavefront, queue modes of the pool kernel), 4 frames in flight
through b200r_pipeline, rasteriser mode 6 + MLAA (TMA-staged strips), wireframe, points; each checked against the blocking call or
the CPU oracle where cheap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import renderer_b200 as rb
from oracle import pyport

W, H = 320, 184
model = pyport.model_path("chessboard.tri")
scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
g = rb.Renderer(0); g.upload(scene)
cams = rb.Orbit.cameras(range(8))
c2 = [rb.make_frame(rb.MODE_RAYTRACE, W, H, cams[k], flags=1 | 4, frame_index=k) for k in range(6)]
a = g.render(c2[0])
assert int((a != pyport.render(scene, c2[0])).sum()) == 0
gen = rb.make_frame(rb.MODE_RAYTRACE, W, H, cams[1], flags=1 | 2 | 4 | 8, ao_samples=4, frame_index=1)
b = g.render(gen)
assert int((b != pyport.render(scene, gen)).sum()) == 0
want = [g.render(f).copy() for f in c2]
pipe = rb.Pipeline(g, W, H, depth=4)
import ctypes as C
hosts = []
for f in c2:
    p = C.c_void_p(); rb.lib().b200r_host_alloc(W * H * 4, C.byref(p)); hosts.append(p)
    pipe.submit(f, p.value)
pipe.drain()
for k, p in enumerate(hosts):
    got = np.ctypeslib.as_array((C.c_uint32 * (W * H)).from_address(p.value)).reshape(H, W)
    assert np.array_equal(got, want[k]), k
pipe.close()
for m, fl in ((6, rb.F_DEFAULT | rb.F_MLAA), (3, rb.F_DEFAULT), (2, rb.F_DEFAULT)):
    f = rb.make_frame(m, W, H, cams[2], flags=fl)
    assert int((g.render(f) != pyport.render(scene, f)).sum()) == 0, m
print("sanitize.py: all paths ran, frames equal the oracle / the blocking call")
g.close()
