"""CPU: host plumbing (loader, BVH, cache, orbit, lights) against what the reference produced."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
INDEX = json.load(open(os.path.join(HERE, "golden", "index.json")))


@pytest.mark.parametrize("model", sorted(INDEX["_bvh_sha256"]))
def test_bvh_is_byte_identical_to_reference_cache(rb, pyport, model):
    """Scene::load + CreateBVH + CreateCFBVH: our .bvh bytes == the .bvh the reference wrote for the model."""
    path = pyport.model_path(model)
    if not os.path.exists(path):
        pytest.skip("model not staged")
    s = rb.Scene(path).UpdateBoundingVolumeHierarchy(None, forceRecalc=True)
    assert hashlib.sha256(s.bvh_bytes()).hexdigest() == INDEX["_bvh_sha256"][model]
    assert 0 <= s.bvh_depth < 32


@pytest.mark.parametrize("model", ["single.ply", "square.ply", "box-and-plane.ply", "torus.ply", "sphere.ply", "x-wing.ply", "tie.ply",
                                   "trainColor.tri", "dragon_vis.ply"])
def test_device_bvh_build_steps_reproduce_the_reference_tree(rb, pyport, model):
    """csrc/bvh_steps.h - the per-item functions the CUDA build kernels are made of - run level by level in plain loops:
    the resulting .bvh bytes equal the cache the reference wrote (and therefore the recursive host builder's)."""
    path = pyport.model_path(model)
    if not os.path.exists(path):
        pytest.skip("model not staged")
    s = rb.Scene(path)
    got, depth = s.bvh_bytes_from_steps_on_host()
    assert hashlib.sha256(got).hexdigest() == INDEX["_bvh_sha256"][model]
    assert 0 <= depth < 32


def test_bvh_cache_roundtrip_and_interop(rb, pyport, tmp_path):
    path = pyport.model_path("torus.ply")
    if not os.path.exists(path):
        pytest.skip("model not staged")
    cache = str(tmp_path / "torus.ply.bvh")
    a = rb.Scene(path).UpdateBoundingVolumeHierarchy(cache)
    assert hashlib.sha256(open(cache, "rb").read()).hexdigest() == INDEX["_bvh_sha256"]["torus.ply"]
    b = rb.Scene(path).UpdateBoundingVolumeHierarchy(cache)          # now read from the cache
    assert a.bvh_bytes() == b.bvh_bytes()
    open(cache, "wb").write(b"\x01\x02\x03")                          # short/corrupt cache -> silent rebuild
    c = rb.Scene(path).UpdateBoundingVolumeHierarchy(cache)
    assert c.bvh_bytes() == a.bvh_bytes()


def test_loader_errors_like_the_reference(rb, tmp_path):
    with pytest.raises(rb.RendererError, match="not found|Missing"):
        rb.Scene(str(tmp_path / "nope.tri"))
    p = tmp_path / "x.obj"; p.write_text("hi")
    with pytest.raises(rb.RendererError, match="extension"):
        rb.Scene(str(p))
    q = tmp_path / "bad.tri"; q.write_bytes(b"\xde\xc0\xad\xde" + b"\x05\x00\x00\x00" + b"\x00" * 7)
    with pytest.raises(rb.RendererError, match="Malformed"):
        rb.Scene(str(q))


def test_tiny_ply_and_counts(rb, pyport):
    path = pyport.model_path("single.ply")
    if not os.path.exists(path):
        pytest.skip("model not staged")
    s = rb.Scene(path).UpdateBoundingVolumeHierarchy()
    assert s.n_triangles == 1 and s.n_nodes == 1 and s.bvh_depth == 0
    c = rb.Scene(pyport.model_path("chessboard.tri"))
    assert (c.n_vertices, c.n_triangles) == (32488, 46658)


def test_orbit_is_the_reference_recurrence(rb):
    cams = rb.Orbit.cameras([0, 1, 99])
    e0 = np.array(cams[0].eye[:])
    assert abs(np.linalg.norm(e0) - 4.8) < 1e-5 and e0[2] == 0.0 and e0[1] < 0      # angle1 = -0.3 deg
    e99 = np.array(cams[99].eye[:])
    assert abs(np.degrees(np.arctan2(-e99[1], e99[0])) - 30.0) < 1e-3               # 100 steps of 0.3 deg
    mv = np.array(cams[0].mv[:]).reshape(3, 3)
    assert np.allclose(mv @ mv.T, np.eye(3), atol=1e-6)
    assert np.allclose(mv[2], -e0 / np.linalg.norm(e0), atol=1e-6)                  # forward looks at the origin


def test_default_lights(rb):
    l0, l1 = rb.default_light_pos(0), rb.default_light_pos(1)
    assert np.allclose(l0, (3.394113, 3.394113, 4.8), atol=1e-5)
    assert np.allclose(l1, (4.8, -4.8, 4.8), atol=1e-6)


@pytest.mark.parametrize("model,mode,size,frame", [("statue.ply", 6, (320, 240), 0), ("statue.ply", 5, (640, 480), 37),
                                                   ("chessboard.tri", 6, (400, 304), 9), ("torus.ply", 4, (64, 48), 3),
                                                   ("dragon_vis.ply", 6, (1280, 720), 1)])
def test_mlaa_step_functions_equal_the_oracle(rb, pyport, model, mode, size, frame):
    """csrc/mlaa_steps.h - the per-item functions the MLAA kernels are made of (flags, line bounds, split heights, the in-place
    blends; with and without the 8-pixel batched loads the device uses) - run in plain loops in the kernels' job order:
    the filtered frame equals the restatement of the reference's MLAA (itself pinned to the reference's goldens) bit for bit."""
    import ctypes as C
    path = pyport.model_path(model)
    if not os.path.exists(path):
        pytest.skip("model not staged")
    s = rb.Scene(path)
    cam = rb.Orbit.cameras([frame])[frame]
    plain = pyport.render(s, rb.make_frame(mode, size[0], size[1], cam))
    want = pyport.mlaa(plain)
    assert 0 < int((want != plain).sum())
    for batched in (0, 1):
        got = np.ascontiguousarray(plain, dtype=np.uint32).copy()
        rc = rb.lib().b200r_selftest_mlaa_steps_host(got.ctypes.data, size[0], size[1], batched)
        assert rc == 0
        assert np.array_equal(got, want), f"batched={batched}: {int((got != want).sum())} pixels differ"


def test_mlaa_step_functions_on_noise(rb, pyport):
    """Random blocky frames (every kind of L/Z/U shape, lines touching all four borders, runs up to a full row) through the
    same comparison: batched == unbatched == oracle."""
    rng = np.random.default_rng(7)
    for W, H, cell in ((64, 48, 1), (128, 96, 3), (256, 64, 8), (96, 256, 5), (512, 384, 16)):
        coarse = rng.integers(0, 4, size=((H + cell - 1) // cell, (W + cell - 1) // cell), dtype=np.uint32)
        palette = np.array([0x00101010, 0x00E0E0E0, 0x00FF2040, 0x0020C0FF], dtype=np.uint32)
        img = np.kron(palette[coarse], np.ones((cell, cell), dtype=np.uint32))[:H, :W].astype(np.uint32)
        img[H // 3, :] = 0x00FFFFFF                    # a separation line as long as the row
        img[:, W // 5] = 0x00000000                    # ... and one as long as the column
        want = pyport.mlaa(img)
        for batched in (0, 1):
            got = np.ascontiguousarray(img).copy()
            assert rb.lib().b200r_selftest_mlaa_steps_host(got.ctypes.data, W, H, batched) == 0
            assert np.array_equal(got, want), f"{W}x{H} cell {cell} batched={batched}: {int((got != want).sum())} pixels differ"


@pytest.mark.parametrize("width", [8, 64, 800, 3840])
def test_span_walkers_equal_the_reference_loop(rb, width):
    """csrc/raster_steps.h: walk_span and the batched walk_span_keyed (8 depth keys requested together, what the device's
    resolve passes run per span) visit the pixels of Screen::RasterizeTriangle's loop with the same interpolant bits -
    random spans incl. single-point, zero-length, clipped left / right / both."""
    import ctypes as C
    bad = C.c_uint64(123)
    assert rb.lib().b200r_selftest_span_walk_host(width + 1, 40000, width, C.byref(bad)) == 0
    assert bad.value == 0
