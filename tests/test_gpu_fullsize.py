"""GPU parity at BASELINE.json's FULL sizes, directly against the unmodified reference: the SHA-256 of the CUDA frame equals
the SHA-256 of the frame the reference presented (tests/golden/index.json, `full_*` cases: frames 0, 1, 37, 99 of the orbit)."""
import os

import pytest

from test_cpu_oracle import FULL_CASES, case_frames, sha256_of

pytestmark = pytest.mark.gpu

# C3 / C5 (ambient occlusion x16 at 1920x1080 / 3840x2160) were added when round 1's GPU budget was spent: their restatement is
# pinned to the reference at full size on the CPU (test_cpu_oracle.py) and the CUDA path to the restatement at 320x240 (goldens),
# but these two have not been run on a GPU yet - B200R_FULLSIZE_GPU=1 enables them.
NOT_YET_RUN_ON_A_GPU = ("full_c3_", "full_c5_")


@pytest.mark.parametrize("name", sorted(FULL_CASES))
def test_cuda_frame_has_the_reference_digest(rb, pyport, load_scene, gpu, name):
    if name.startswith(NOT_YET_RUN_ON_A_GPU) and not os.environ.get("B200R_FULLSIZE_GPU"):
        pytest.skip("full-size AO case not yet run on a GPU (set B200R_FULLSIZE_GPU=1)")
    c, frames = case_frames(rb, name)
    s = load_scene(c["model"], bvh=c["mode"] >= 9)
    gpu.upload(s)
    for k, want_sha, f in frames:
        got = gpu.render(f)
        if sha256_of(got) != want_sha:
            from util import diff_stats
            n, mx, over = diff_stats(got, pyport.render(s, f))       # the restatement has the reference's digest: diff against it
            raise AssertionError(f"{name} frame {k}: {n} pixels differ from the reference, max channel delta {mx}, "
                                 f"{over} beyond 1 LSB")
