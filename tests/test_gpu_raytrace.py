"""GPU parity: the CUDA ray tracer (through the C-ABI) against the CPU restatement on the same inputs.

Mirrors how one would test the reference: load a model, follow the `-b` orbit, render mode 9 / 0."""
import pytest

from util import assert_parity

pytestmark = pytest.mark.gpu

W, H = 320, 240


@pytest.mark.parametrize("model,frames", [("torus.ply", [0, 37]), ("chessboard.tri", [0, 99]),
                                          ("dragon_vis.ply", [1]), ("trainColor.tri", [5]), ("single.ply", [0])])
def test_raytrace_default(rb, pyport, load_scene, gpu, model, frames):
    s = load_scene(model)
    gpu.upload(s)
    for k, cam in rb.Orbit.cameras(frames).items():
        f = rb.make_frame(rb.MODE_RAYTRACE, W, H, cam, frame_index=k)
        assert_parity(gpu.render(f), pyport.render(s, f), f"{model} frame {k} mode 9")


@pytest.mark.parametrize("flags", [0, 1, 2, 4, 3, 5, 6])   # all subsets of shadows/reflections/phong but the default
def test_raytrace_feature_switches(rb, pyport, load_scene, gpu, flags):
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    cam = rb.Orbit.cameras([12])[12]
    f = rb.make_frame(rb.MODE_RAYTRACE, W, H, cam, flags=flags)
    assert_parity(gpu.render(f), pyport.render(s, f), f"chessboard flags={flags}")


def test_raytrace_antialias_two_lights(rb, pyport, load_scene, gpu):
    s = load_scene("torus.ply")
    gpu.upload(s)
    cam = rb.Orbit.cameras([3])[3]
    f = rb.make_frame(rb.MODE_RAYTRACE_AA, W, H, cam, n_lights=2)
    assert_parity(gpu.render(f), pyport.render(s, f), "torus AA 2 lights")


@pytest.mark.parametrize("model", ["torus.ply", "chessboard.tri"])
def test_raytrace_ambient_occlusion(rb, pyport, load_scene, gpu, model):
    s = load_scene(model)
    gpu.upload(s)
    cam = rb.Orbit.cameras([2])[2]
    f = rb.make_frame(rb.MODE_RAYTRACE, 160, 120, cam, flags=rb.F_DEFAULT | rb.F_AO, ao_samples=16, frame_index=2)
    assert_parity(gpu.render(f), pyport.render(s, f), f"{model} AO x16")


def test_counters_match_oracle(rb, pyport, load_scene, gpu):
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    cam = rb.Orbit.cameras([0])[0]
    f = rb.make_frame(rb.MODE_RAYTRACE, W, H, cam, flags=rb.F_SHADOWS | rb.F_PHONG_NORMAL)
    gpu.set_counters(True)
    try:
        img = gpu.render(f)
        got = gpu.counters()
    finally:
        gpu.set_counters(False)
    want_img, want = pyport.render(s, f, counters=True)
    assert_parity(img, want_img, "chessboard counters frame")
    for k in ("rays_primary", "rays_shadow", "rays_reflection", "rays_ao", "node_tests", "leaf_visits", "tri_tests"):
        assert got[k] == want[k], (k, got[k], want[k])


def test_row_sharding_matches_full_frame(rb, load_scene, gpu):
    import numpy as np
    s = load_scene("torus.ply")
    gpu.upload(s)
    cam = rb.Orbit.cameras([7])[7]
    full = gpu.render(rb.make_frame(rb.MODE_RAYTRACE, W, H, cam))
    P = 4
    for r in range(P):
        part = gpu.render(rb.make_frame(rb.MODE_RAYTRACE, W, H, cam, row_first=r, row_step=P))
        assert np.array_equal(part, full[r::P])


def test_split_pipeline_equals_monolithic_kernel(rb, load_scene, gpu):
    """Mode 9 runs as root-cull -> persistent primary traversal -> shade; the single persistent kernel (used for
    mode 0) must give the same frame and the same work counters."""
    import numpy as np
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    cam = rb.Orbit.cameras([40])[40]
    f = rb.make_frame(rb.MODE_RAYTRACE, 640, 360, cam)
    gpu.set_counters(True)
    try:
        a = gpu.render(f); ca = gpu.counters()
        with gpu.switch("monolithic_rt"):
            b = gpu.render(f); cb = gpu.counters()
    finally:
        gpu.set_counters(False)
    assert np.array_equal(a, b)
    assert ca == cb


@pytest.mark.parametrize("model,size", [("chessboard.tri", (1920, 1080)), ("dragon_vis.ply", (1280, 720)), ("tie.ply", (1280, 720)),
                                        ("x-wing.ply", (800, 600)), ("kerolamp.ply", (800, 600))])
def test_distance_pruning_changes_nothing(rb, pyport, load_scene, gpu, model, size):
    """Near-first ordering + conservative distance pruning of the primary-ray kernel vs the reference's full traversal:
    same frame, on scenes that do (chessboard, tie, x-wing, kerolamp) and do not contain triangles flagged unprunable."""
    import numpy as np
    s = load_scene(model)
    gpu.upload(s)
    for k, cam in rb.Orbit.cameras([0, 57]).items():
        f = rb.make_frame(rb.MODE_RAYTRACE, size[0], size[1], cam)
        pruned = gpu.render(f)
        with gpu.switch("no_prune"):
            plain = gpu.render(f)
        assert np.array_equal(pruned, plain), f"{model} frame {k}"
        if k == 0:
            assert_parity(pruned, pyport.render(s, f), f"{model} {size} frame {k}")


@pytest.mark.parametrize("flags", [1 | 4, 4, 1, 0])
@pytest.mark.parametrize("model", ["chessboard.tri", "dragon_vis.ply", "trainColor.tri"])
def test_fused_shadow_continuation_equals_hit_record_path(rb, pyport, load_scene, gpu, model, flags):
    """One light, no reflections, no AO: rt_pool_kernel shades a resolved hit itself and re-arms the slot as the hit's shadow ray.
    Must equal the generic route (hit records + rt_shade_kernel), round 1's job pipeline, and the oracle."""
    import numpy as np
    s = load_scene(model)
    gpu.upload(s)
    cam = rb.Orbit.cameras([33])[33]
    f = rb.make_frame(rb.MODE_RAYTRACE, 1280, 720, cam, flags=flags)
    fused = gpu.render(f)
    with gpu.switch("no_fuse"):
        split = gpu.render(f)
    with gpu.switch("rt_legacy"):
        legacy = gpu.render(f)
    assert np.array_equal(fused, split) and np.array_equal(legacy, split)
    assert_parity(fused, pyport.render(s, f), f"{model} flags={flags} fused shadow rays")


@pytest.mark.parametrize("flags", [1 | 4, 1 | 2 | 4, 4])
@pytest.mark.parametrize("model,size", [("chessboard.tri", (1920, 1080)), ("dragon_vis.ply", (1280, 720)), ("single.ply", (320, 240)),
                                        ("trainColor.tri", (800, 600)), ("torus.ply", (640, 480))])
def test_pool_overflow_guard_and_legacy_pipeline_agree(rb, pyport, load_scene, gpu, model, size, flags):
    """rt_pool_kernel with 128-entry pools (the overflow guard - lanes walking whole subtrees with a private stack - runs all
    the time) and round 1's lane-per-job pipeline against the default 512-entry pools: identical frames, fused and generic
    configurations, and equal to the oracle."""
    import numpy as np
    s = load_scene(model)
    gpu.upload(s)
    for k, cam in rb.Orbit.cameras([3, 64]).items():
        f = rb.make_frame(rb.MODE_RAYTRACE, size[0], size[1], cam, flags=flags)
        pool = gpu.render(f)
        with gpu.switch("pool_small"):
            small = gpu.render(f)
        assert np.array_equal(small, pool), f"{model} frame {k} flags={flags}: 128-entry pools"
        with gpu.switch("rt_legacy"):
            legacy = gpu.render(f)
        assert np.array_equal(legacy, pool), f"{model} frame {k} flags={flags}: job pipeline"
        with gpu.switch("no_wavefront"):
            threads = gpu.render(f)
        assert np.array_equal(threads, pool), f"{model} frame {k} flags={flags}: rt_shade_kernel instead of the wavefront"
        if k == 3 and size[0] <= 1280:
            assert_parity(pool, pyport.render(s, f), f"{model} {size} frame {k} flags={flags}")


@pytest.mark.parametrize("model,size", [("chessboard.tri", (1920, 1080)), ("dragon_vis.ply", (801, 603)), ("torus.ply", (64, 48))])
def test_pool_scheduling_variants_change_nothing(rb, load_scene, gpu, model, size):
    """How rt_pool_kernel deals pixels to warps (scattered 4-pixel groups / whole tiles) and pops its pool (policies 0, 1, 2) is
    scheduling only: every combination must give the same frame, fused and generic configurations."""
    import numpy as np
    s = load_scene(model)
    gpu.upload(s)
    cam = rb.Orbit.cameras([21])[21]
    for flags in (1 | 4, 1 | 2 | 4):
        f = rb.make_frame(rb.MODE_RAYTRACE, size[0], size[1], cam, flags=flags)
        base = gpu.render(f)
        for scatter in (1, 2):
            for policy in (0, 1, 2, 3):
                for occ3 in (0, 1):
                    gpu.set_switch("pool_scatter", scatter); gpu.set_switch("pool_policy", policy); gpu.set_switch("pool_occ3", occ3)
                    try:
                        got = gpu.render(f)
                    finally:
                        gpu.set_switch("pool_scatter", 0); gpu.set_switch("pool_policy", 0); gpu.set_switch("pool_occ3", 0)
                    assert np.array_equal(got, base), f"{model} flags={flags} scatter={scatter} policy={policy} occ3={occ3}"
        # CTA size (default 4 independent warps per CTA; 8, 2, 1: the same number of warps per SM): scheduling only, too
        for warps in (8, 2, 1):
            for scatter in (1, 2):
                gpu.set_switch("pool_cta_warps", warps); gpu.set_switch("pool_scatter", scatter)
                try:
                    got = gpu.render(f)
                finally:
                    gpu.set_switch("pool_cta_warps", 0); gpu.set_switch("pool_scatter", 0)
                assert np.array_equal(got, base), f"{model} flags={flags} cta_warps={warps} scatter={scatter}"
    if size[0] <= 801:      # AO + reflections: the queue modes of the kernel with small CTAs
        f = rb.make_frame(rb.MODE_RAYTRACE, size[0], size[1], cam, flags=1 | 2 | 4 | 8, ao_samples=4)
        base = gpu.render(f)
        for warps in (8, 2):
            with gpu.switch("pool_cta_warps", warps):
                got = gpu.render(f)
            assert np.array_equal(got, base), f"{model} AO cta_warps={warps}"


@pytest.mark.parametrize("model", ["chessboard.tri", "dragon_vis.ply", "trainColor.tri"])
def test_root_box_screen_rectangle_culls_nothing_visible(rb, load_scene, gpu, model):
    """K0 skips ray construction for pixels outside a conservative screen rectangle of the root box; the frame must be
    identical to the one where every pixel's ray takes the root test (several orbit positions, two aspect ratios)."""
    import numpy as np
    s = load_scene(model)
    gpu.upload(s)
    cams = rb.Orbit.cameras([0, 17, 40, 77])
    for k, cam in cams.items():
        for (w, h) in ((640, 360), (320, 480)):
            f = rb.make_frame(rb.MODE_RAYTRACE, w, h, cam, flags=1 | 4)
            culled = gpu.render(f)
            with gpu.switch("no_root_rect"):
                full = gpu.render(f)
            assert np.array_equal(culled, full), f"{model} frame {k} {w}x{h}"


def test_pipelined_host_render_equals_blocking_call(rb, load_scene, gpu):
    """b200r_render_async/b200r_wait (two frames in flight, copy-out overlapped) delivers the same frames as b200r_render,
    into page-locked and into pageable caller memory."""
    import numpy as np
    import torch
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    cams = rb.Orbit.cameras(range(6))
    frames = [rb.make_frame(rb.MODE_RAYTRACE, 640, 360, cams[k], flags=1 | 4, frame_index=k) for k in range(6)]
    want = [gpu.render(f).copy() for f in frames]
    pinned = [torch.zeros((360, 640), dtype=torch.int32).pin_memory() for _ in range(6)]
    outs = [p.numpy().view(np.uint32) for p in pinned]
    for f, o in zip(frames, outs):
        gpu.render_async(f, o)
    gpu.wait()
    for k in range(6):
        assert np.array_equal(outs[k], want[k]), f"pinned frame {k}"
    pageable = [np.zeros((360, 640), dtype=np.uint32) for _ in range(6)]
    for f, o in zip(frames, pageable):
        gpu.render_async(f, o)
    # a blocking call drains the pipeline first
    again = gpu.render(frames[0])
    assert np.array_equal(again, want[0])
    for k in range(6):
        assert np.array_equal(pageable[k], want[k]), f"pageable frame {k}"


def test_overlapped_async_frames_equal_blocking_frames(rb, load_scene, gpu):
    """b200r_render_async alternates ray-traced frames between two streams / scratch sets so that frame i+1 fills the SMs the
    tail of frame i leaves idle. Frames of different sizes, modes and feature sets submitted back to back must each equal
    the blocking call's frame."""
    import numpy as np
    import torch
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    cams = rb.Orbit.cameras(range(12))
    specs = [(rb.MODE_RAYTRACE, 1920, 1080, 1 | 4), (rb.MODE_RAYTRACE, 1920, 1080, 1 | 4), (rb.MODE_RAYTRACE, 1280, 720, 1 | 2 | 4),
             (rb.MODE_PHONG, 800, 600, 0), (rb.MODE_RAYTRACE, 1920, 1080, 1 | 4), (rb.MODE_RAYTRACE_AA, 320, 240, 1 | 2 | 4),
             (rb.MODE_RAYTRACE, 1920, 1080, 4), (rb.MODE_RAYTRACE, 1920, 1080, 1 | 4), (rb.MODE_RAYTRACE, 640, 360, 1 | 4),
             (rb.MODE_RAYTRACE, 1920, 1080, 1 | 4), (rb.MODE_RAYTRACE, 1920, 1080, 1 | 4), (rb.MODE_RAYTRACE, 1920, 1080, 1 | 4)]
    frames = [rb.make_frame(m, w, h, cams[k], flags=fl, frame_index=k) for k, (m, w, h, fl) in enumerate(specs)]
    want = [gpu.render(f).copy() for f in frames]
    for rep in range(2):
        pinned = [torch.zeros((h, w), dtype=torch.int32).pin_memory() for (_, w, h, _) in specs]
        outs = [p.numpy().view(np.uint32) for p in pinned]
        for f, o in zip(frames, outs):
            gpu.render_async(f, o)
        gpu.wait()
        for k in range(len(frames)):
            assert np.array_equal(outs[k], want[k]), f"rep {rep} frame {k} {specs[k]}"


def test_frames_in_flight_on_scratch_slots_equal_blocking_frames(rb, load_scene, gpu):
    """b200r_pipeline (world 1): frames rotating over 4 streams / scratch sets, generic (AO + reflections, the wavefront) and
    C2-type frames, and b200r_render_async at pipeline depth 4, deliver the blocking call's frames bit for bit."""
    import numpy as np
    import torch
    s = load_scene("chessboard.tri")
    gpu.upload(s)
    W, H, n = 1280, 720, 8
    cams = rb.Orbit.cameras(range(n))
    frames = [rb.make_frame(rb.MODE_RAYTRACE, W, H, cams[k], flags=(1 | 4) if k % 2 == 0 else (1 | 2 | 4 | 8), ao_samples=8, frame_index=k)
              for k in range(n)]
    want = [gpu.render(f).copy() for f in frames]
    pipe = rb.Pipeline(gpu, W, H, depth=4)
    try:
        hosts = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(n)]
        for base in (0, 4):
            for k in range(4):
                pipe.submit(frames[base + k], hosts[base + k].data_ptr())
            pipe.drain()
            for k in range(4):
                assert np.array_equal(hosts[base + k].numpy().view(np.uint32), want[base + k]), f"pipeline frame {base + k}"
        # a slot out of range / the wireframe mode are refused, not rendered
        import pytest
        st = torch.cuda.Stream()
        with pytest.raises(Exception):
            gpu.render_device_slot(frames[0], pipe.slot_frame(0), st.cuda_stream, rb.MAX_FRAMES_IN_FLIGHT)
        with pytest.raises(Exception):
            gpu.render_device_slot(rb.make_frame(rb.MODE_LINES, W, H, cams[0]), pipe.slot_frame(0), st.cuda_stream, 0)
    finally:
        pipe.close()
    gpu.set_pipeline_depth(4)
    pinned = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(n)]
    outs = [p.numpy().view(np.uint32) for p in pinned]
    for rep in range(2):
        for f, o in zip(frames, outs):
            gpu.render_async(f, o)
        gpu.wait()
        for k in range(n):
            assert np.array_equal(outs[k], want[k]), f"depth-4 async frame {k} rep {rep}"
    gpu.set_pipeline_depth(2)
