"""CPU: bench.py's contract where no GPU is needed - the reference arm (the unmodified reference timed on the host cores), the
conversion of its fps into the headline metric, the roofline's byte model, and that the B200 arm has no CPU fallback."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_rays_per_frame_fixture_is_exact_for_c2_and_interpolated_elsewhere():
    import bench
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "rays_per_frame.json")))
    c2 = {int(k): v for k, v in d["c2"].items()}
    assert set(range(0, 128)) <= set(c2)                                  # every frame the default bench times (5..104)
    want = sum(c2[k] for k in range(5, 105)) / 100.0
    assert bench.fixture_rays_per_frame("c2", list(range(5, 105))) == pytest.approx(want, rel=0, abs=1e-6)
    c3 = {int(k): v for k, v in d["c3"].items()}
    lo, hi = c3[16], c3[24]
    assert bench.fixture_rays_per_frame("c3", [20]) == pytest.approx((lo + hi) / 2.0)
    assert bench.fixture_rays_per_frame("c3", [100000]) == c3[max(c3)]    # past the last stored frame: clamped


def test_algorithmic_bytes_follow_survey_8d():
    import bench
    c = dict(node_tests=10, leaf_visits=5, tri_tests=7, tris_setup=3, z_tests=11, z_passes=4)
    assert bench.algorithmic_bytes(c, 100, 10) == 32 * 15 + 84 * 7 + 4 * 100 * 10
    assert bench.algorithmic_bytes(c, 100, 10, raster=True) == (28 * 3 + 144) * 3 + 8 * 11 + 4 * 4 + 8 * 100 * 10


def test_reference_arm_line_and_idle_ranks():
    from oracle import pyport
    import bench
    wl = bench.WORKLOADS["c2"]
    if not os.path.exists(pyport.ref_exe(wl["W"], wl["H"], no_reflections=True, fast=True)):
        pytest.skip("oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, B200R_REF_STEP_SECONDS="0.4")
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["config"]["workload"] == wl["desc"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] == pytest.approx(d["config"]["rays_per_frame"] * d["fps"] / 1e6)
    # under torchrun only rank 0 works: the others exit 0 and print nothing
    env2 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"],
                        cwd=ROOT, env=env2, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
