"""GPU parity: the CUDA rasteriser (modes 1, 2, 4..8 + shadow-map pre-pass) through the C-ABI against the CPU
restatement, and against the golden frames rendered by the unmodified reference (tests/golden/)."""
import numpy as np
import pytest

from test_cpu_oracle import CASES, case_frames
from util import assert_parity

pytestmark = pytest.mark.gpu

W, H = 320, 240


def setup_shadowmaps(rb, gpu, scene, frame):
    """GPU-rendered shadow maps for the frame's lights; returns them (downloaded) for the oracle."""
    maps = []
    for i in range(frame.n_lights):
        gpu.render_shadowmap(i, tuple(frame.lights[i].pos))
        maps.append(gpu.download_shadowmap(i))
    return tuple(maps)


@pytest.mark.parametrize("model", ["statue.ply", "chessboard.tri", "torus.ply"])
def test_shadowmap_matches_oracle_bit_for_bit(rb, pyport, load_scene, gpu, model):
    """Light::RenderSceneIntoShadowBuffer on the device == the restatement, all 1024x1024 floats."""
    s = load_scene(model, bvh=False)
    gpu.upload(s)
    for li in (0, 1):
        lp = rb.default_light_pos(li)
        gpu.render_shadowmap(li, lp)
        got = gpu.download_shadowmap(li)
        want = pyport.shadowmaps_for(s, rb.make_frame(7, W, H, rb.Camera(), n_lights=2))[li]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{model} light {li}"


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("model,frame_no,lights", [("statue.ply", 0, 1), ("chessboard.tri", 31, 2), ("trainColor.tri", 5, 1),
                                                   ("dragon_vis.ply", 77, 1)])
def test_raster_modes_vs_oracle(rb, pyport, load_scene, gpu, mode, model, frame_no, lights):
    s = load_scene(model, bvh=False)
    gpu.upload(s)
    cam = rb.Orbit.cameras([frame_no])[frame_no]
    f = rb.make_frame(mode, W, H, cam, n_lights=lights)
    maps = setup_shadowmaps(rb, gpu, s, f) if mode in (7, 8) else ()
    want = pyport.render(s, f, shadowmaps=maps)
    assert_parity(gpu.render(f), want, f"{model} mode {mode} frame {frame_no}")


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items()
                                        if c["mode"] in (1, 2, 3, 4, 5, 6, 7, 8) and not c["variant"].get("mlaa")))
def test_raster_vs_reference_golden(rb, load_scene, gpu, name):
    c, frames = case_frames(rb, name)
    s = load_scene(c["model"], bvh=False)
    gpu.upload(s)
    for k, want, f in frames:
        if c["mode"] in (7, 8):
            setup_shadowmaps(rb, gpu, s, f)
        assert_parity(gpu.render(f), want, f"{name} frame {k}")


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["mode"] in (0, 9) and not c["variant"].get("mlaa")))
def test_raytrace_vs_reference_golden(rb, load_scene, gpu, name):
    c, frames = case_frames(rb, name)
    s = load_scene(c["model"])
    gpu.upload(s)
    for k, want, f in frames:
        assert_parity(gpu.render(f), want, f"{name} frame {k}")


def test_raster_counters_and_row_sharding(rb, pyport, load_scene, gpu):
    s = load_scene("statue.ply", bvh=False)
    gpu.upload(s)
    cam = rb.Orbit.cameras([0])[0]
    f = rb.make_frame(6, 800, 600, cam)
    gpu.set_counters(True)
    try:
        full = gpu.render(f)
        got = gpu.counters()
    finally:
        gpu.set_counters(False)
    want_img, want = pyport.render(s, f, counters=True)
    assert_parity(full, want_img, "statue 800x600 mode 6")
    assert (got["tris_setup"], got["z_tests"]) == (want["tris_setup"], want["z_tests"])
    assert (want["tris_setup"], want["z_tests"]) == (26163, 304982)          # SURVEY.md §8d probe of the reference
    for r in range(3):
        part = gpu.render(rb.make_frame(6, 800, 600, cam, row_first=r, row_step=3))
        assert np.array_equal(part, full[r::3])


def test_full_size_c4_properties(rb, pyport, load_scene, gpu):
    """BASELINE config C4 (statue.ply 3840x2160, modes 5 and 6) at full size: bit-exact against the restatement,
    plus size-independent properties (idempotence, black background share)."""
    s = load_scene("statue.ply", bvh=False)
    gpu.upload(s)
    cam = rb.Orbit.cameras([3])[3]
    for mode in (5, 6):
        f = rb.make_frame(mode, 3840, 2160, cam)
        a = gpu.render(f)
        b = gpu.render(f)
        assert np.array_equal(a, b)
        assert_parity(a, pyport.render(s, f), f"statue 4K mode {mode}")
        assert 0.02 < float((a != 0).mean()) < 0.6


@pytest.mark.parametrize("model,size", [("chessboard.tri", (1920, 1080)), ("statue.ply", (1280, 720)), ("torus.ply", (3840, 2160))])
def test_wireframe_full_size(rb, pyport, load_scene, gpu, model, size):
    """Mode 3 must be BIT-exact (integer Wu lines + ordered alpha blending), including lines clipped at the borders."""
    import numpy as np
    s = load_scene(model, bvh=False)
    gpu.upload(s)
    cam = rb.Orbit.cameras([21])[21]
    f = rb.make_frame(3, size[0], size[1], cam)
    got = gpu.render(f)
    assert np.array_equal(got, pyport.render(s, f)), f"{model} {size}"
    for r in range(2):
        part = gpu.render(rb.make_frame(3, size[0], size[1], cam, row_first=r, row_step=2))
        assert np.array_equal(part, got[r::2])


def test_rasterised_frames_in_flight_equal_blocking_frames(rb, load_scene, gpu):
    """Raster modes rotate over scratch sets / streams like ray-traced frames do (b200r_render_async, b200r_pipeline): nothing is
    read back inside a frame any more. Frames of several modes - with and without MLAA - submitted back to back must each
    equal the blocking call's frame, through render_async and through a 3-deep pipeline."""
    import numpy as np
    import torch
    s = load_scene("statue.ply")
    gpu.upload(s)
    gpu.render_shadowmap(0, rb.default_light_pos(0))
    cams = rb.Orbit.cameras(range(10))
    specs = [(6, 1280, 720, rb.F_DEFAULT | rb.F_MLAA), (5, 1280, 720, rb.F_DEFAULT), (6, 1280, 720, rb.F_DEFAULT | rb.F_MLAA),
             (4, 800, 600, rb.F_DEFAULT), (8, 1280, 720, rb.F_DEFAULT | rb.F_MLAA), (2, 640, 480, rb.F_DEFAULT),
             (6, 1280, 720, rb.F_DEFAULT | rb.F_MLAA), (7, 1280, 720, rb.F_DEFAULT), (1, 640, 480, rb.F_DEFAULT),
             (6, 1280, 720, rb.F_DEFAULT | rb.F_MLAA)]
    frames = [rb.make_frame(m, w, h, cams[k], flags=fl, frame_index=k) for k, (m, w, h, fl) in enumerate(specs)]
    want = [gpu.render(f).copy() for f in frames]
    for depth in (2, 3):
        gpu.set_pipeline_depth(depth)
        outs = [torch.zeros((f.height, f.width), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for f in frames]
        for f, o in zip(frames, outs):
            gpu.render_async(f, o)
        gpu.wait()
        for k in range(len(frames)):
            assert np.array_equal(outs[k], want[k]), f"render_async depth {depth}, frame {k} (mode {specs[k][0]})"
    gpu.set_pipeline_depth(2)
    same = [k for k, sp in enumerate(specs) if sp[:3] == (6, 1280, 720)]
    pipe = rb.Pipeline(gpu, 1280, 720, depth=3)
    try:
        hosts = [torch.zeros((720, 1280), dtype=torch.int32).pin_memory() for _ in same]
        for k, h in zip(same, hosts):
            pipe.submit(frames[k], h.data_ptr())
        pipe.drain()
        for k, h in zip(same, hosts):
            assert np.array_equal(h.numpy().view(np.uint32), want[k]), f"pipeline frame {k}"
    finally:
        pipe.close()
