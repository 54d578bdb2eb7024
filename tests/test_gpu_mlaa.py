"""GPU parity: the MLAA post filter (reference src/MLAA.cc, hooked in Screen::ShowScreen) on the device."""
import numpy as np
import pytest

from test_cpu_oracle import CASES, case_frames
from test_gpu_raster import setup_shadowmaps
from util import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["variant"].get("mlaa")))
def test_mlaa_vs_reference_golden(rb, load_scene, gpu, name):
    c, frames = case_frames(rb, name)
    mode = 10 if c["mode"] == 0 else c["mode"]
    s = load_scene(c["model"], bvh=mode >= 9)
    gpu.upload(s)
    for k, want, f in frames:
        if mode in (7, 8):
            setup_shadowmaps(rb, gpu, s, f)
        assert f.flags & rb.F_MLAA
        assert_parity(gpu.render(f), want, f"{name} frame {k}")


@pytest.mark.parametrize("model,mode,size", [("statue.ply", 6, (800, 600)), ("chessboard.tri", 9, (640, 480)),
                                             ("statue.ply", 5, (3840, 2160)), ("chessboard.tri", 6, (1920, 1080))])
def test_mlaa_vs_oracle(rb, pyport, load_scene, gpu, model, mode, size):
    s = load_scene(model, bvh=mode >= 9)
    gpu.upload(s)
    cam = rb.Orbit.cameras([9])[9]
    f = rb.make_frame(mode, size[0], size[1], cam, flags=rb.F_DEFAULT | rb.F_MLAA)
    got = gpu.render(f)                                  # default: two-stage blending (all lines first, then the ordered blends)
    want = pyport.render(s, f)
    assert_parity(got, want, f"{model} mode {mode} {size} + MLAA")
    with gpu.switch("mlaa_scan"):
        scan = gpu.render(f)                             # row-scanning kernels: lines found inside the ordered loop
    assert np.array_equal(got, scan)
    with gpu.switch("mlaa_fullscan"):
        full = gpu.render(f)                             # two-stage, but the scanning thread also walks the line it finds
    assert np.array_equal(got, full)
    with gpu.switch("mlaa_no_tma"):
        in_l2 = gpu.render(f)                            # vertical blends walked in L2 instead of a TMA-staged strip in shared memory
    assert np.array_equal(got, in_l2)
    with gpu.switch("mlaa_nobatch"):
        stepwise = gpu.render(f)                         # flag / pixel words loaded one step at a time, as the reference's loops do
    assert np.array_equal(got, stepwise)
    plain = gpu.render(rb.make_frame(mode, size[0], size[1], cam))
    assert 0 < int((plain != got).sum()) < 0.2 * got.size       # the filter touches edges only


def test_mlaa_rejects_sizes_the_reference_cannot_handle(rb, load_scene, gpu):
    s = load_scene("torus.ply", bvh=False)
    gpu.upload(s)
    cam = rb.Orbit.cameras([0])[0]
    with pytest.raises(rb.RendererError, match="MLAA"):
        gpu.render(rb.make_frame(6, 322, 240, cam, flags=rb.F_DEFAULT | rb.F_MLAA))
