#!/bin/bash
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x -k "split_pipeline" 2>&1 | tail -5
compute-sanitizer --tool memcheck python - <<PY 2>&1 | tail -30
import sys; sys.path.insert(0,'.')
import renderer_b200 as rb
from oracle import pyport
p=pyport.model_path('chessboard.tri')
s=rb.Scene(p).UpdateBoundingVolumeHierarchy(p+'.bvh')
g=rb.Renderer(0); g.upload(s)
cam=rb.Orbit.cameras([40])[40]
f=rb.make_frame(9,640,360,cam)
g.set_counters(True)
a=g.render(f); print(g.counters())
PY
