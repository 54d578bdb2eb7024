"""GPU parity at BASELINE.json's FULL sizes, directly against the unmodified reference: the SHA-256 of the CUDA frame equals
the SHA-256 of the frame the reference presented (tests/golden/index.json, `full_*` cases: frames 0, 1, 37, 99 of the orbit)."""
import pytest

from test_cpu_oracle import FULL_CASES, case_frames, sha256_of

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("name", sorted(FULL_CASES))
def test_cuda_frame_has_the_reference_digest(rb, pyport, load_scene, gpu, name):
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
