"""GPU parity: the SAH BVH build as CUDA kernels (b200r_build_bvh) against the .bvh caches the reference wrote."""
import hashlib
import json
import os
import time

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
INDEX = json.load(open(os.path.join(HERE, "golden", "index.json")))


@pytest.mark.parametrize("model", sorted(INDEX["_bvh_sha256"]))
def test_device_bvh_build_is_byte_identical_to_reference_cache(rb, pyport, gpu, model):
    """CreateBVH + CreateCFBVH on the device: nodes + triangle index list == the bytes of the reference's cache file,
    for every model of the reference (1 ... 65 534 triangles, depth 0 ... 20)."""
    path = pyport.model_path(model)
    if not os.path.exists(path):
        pytest.skip("model not staged")
    s = rb.Scene(path)
    got, depth = s.bvh_bytes_from_device_build(gpu)
    assert hashlib.sha256(got).hexdigest() == INDEX["_bvh_sha256"][model]
    assert 0 <= depth < 32


def test_device_built_bvh_renders_the_same_frame(rb, pyport, load_scene, gpu):
    """End to end: upload the device-built tree instead of the host-built one and ray trace."""
    import ctypes as C
    import numpy as np
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    cam = rb.Orbit.cameras([21])[21]
    f = rb.make_frame(rb.MODE_RAYTRACE, 640, 360, cam)
    want = gpu.render(f).copy()
    t0 = time.time()
    got_bytes, _ = s.bvh_bytes_from_device_build(gpu)
    dt = time.time() - t0
    assert got_bytes == s.bvh_bytes()
    print(f"device BVH build of chessboard.tri: {dt * 1e3:.1f} ms wall (second call, incl. copies)")
    assert np.array_equal(gpu.render(f), want)


def test_scene_handle_builds_on_device_and_writes_the_reference_cache(rb, pyport, gpu, tmp_path):
    """b200r_scene_build_bvh_device: no cache -> device build + cache file with the reference's bytes; second call reads it."""
    path = pyport.model_path("trainColor.tri")
    cache = str(tmp_path / "trainColor.tri.bvh")
    a = rb.Scene(path).UpdateBoundingVolumeHierarchyOnDevice(gpu, cache)
    assert hashlib.sha256(open(cache, "rb").read()).hexdigest() == INDEX["_bvh_sha256"]["trainColor.tri"]
    b = rb.Scene(path).UpdateBoundingVolumeHierarchy(cache)           # the host path reads the same cache
    assert a.bvh_bytes() == b.bvh_bytes() and a.bvh_depth == b.bvh_depth
