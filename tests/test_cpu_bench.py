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


def test_b200_arm_assembles_its_line_on_stand_ins(rb, pyport, monkeypatch, capsys):
    """run_b200_arm() with stand-ins for torch.cuda, the device renderer and the frame pipeline (there is no GPU here): every call it
    makes exists with that signature on the stand-ins' real counterparts, the protocol runs in the documented order (all slots warmed,
    K timed submits, e2e on its own pipeline) and the JSON line carries the contract's keys with consistent arithmetic."""
    import types
    import bench
    if not os.path.exists(pyport.model_path("chessboard.tri")):
        pytest.skip("model not staged")
    log = []

    class FakeTensor:
        def __init__(self, pinned=False): self.pinned = pinned
        def data_ptr(self): return 4096
        def zero_(self): log.append("flush"); return self
        def pin_memory(self): return FakeTensor(True)

    class FakeEvent:
        def __init__(self, enable_timing=False): pass
        def record(self, stream=None): pass
        def elapsed_time(self, other): return 0.25

    class FakeStream:
        cuda_stream = 77

    fake_torch = types.ModuleType("torch")
    fake_torch.int32, fake_torch.uint8, fake_torch.float64, fake_torch.int64 = "i32", "u8", "f64", "i64"
    fake_torch.zeros = lambda *a, **k: FakeTensor()
    fake_torch.empty = lambda *a, **k: FakeTensor()
    fake_torch.cuda = types.SimpleNamespace(set_device=lambda d: None, Stream=FakeStream, set_stream=lambda s: None,
                                            synchronize=lambda: None, Event=FakeEvent)
    monkeypatch.setitem(sys.modules, "torch", fake_torch)

    counters = dict(rays_primary=1000, rays_shadow=500, rays_reflection=0, rays_ao=0, node_tests=40000, leaf_visits=5000, tri_tests=9000,
                    tris_setup=0, spans=0, z_tests=0, z_passes=0)

    class FakeRenderer:
        def __init__(self, device): log.append(("renderer", device))
        def upload(self, scene): pass
        def set_counters(self, on): pass
        def render_device(self, frame, ptr, stream): assert isinstance(frame, rb.Frame)
        def counters(self): return dict(counters)
        def last_launches(self): return 1
        def close(self): log.append("renderer.close")
    real_pipeline = rb.Pipeline

    class FakePipeline:
        def __init__(self, renderer, width, height, depth=2, rank=0, world=1, unique_id=None, assemble=rb.ASSEMBLE_PUSH):
            assert 1 <= depth <= rb.MAX_FRAMES_IN_FLIGHT and world == 1 and unique_id is None
            self.depth, self.n, self.flush = depth, 0, 0
            log.append(("pipeline", depth))
        def submit(self, frame, host=None): self.n += 1; log.append(("submit", self.depth, host is not None))
        def drain(self): pass
        def fence(self, stream, pipeline_waits): assert stream == 77
        def set_l2_flush(self, nbytes, prefetch_scene=True): self.flush = nbytes
        def launches(self, reset=False): return 2 * self.n
        def set_timing(self, enabled): pass
        def kernel_ms(self): return (0.6 * 10, 10)
        def close(self): log.append(("pipeline.close", self.depth))
    for name in ("submit", "drain", "fence", "set_l2_flush", "launches", "set_timing", "kernel_ms", "close"):
        assert hasattr(real_pipeline, name), name                # the stand-in does not invent API
    monkeypatch.setattr(rb, "Renderer", FakeRenderer)
    monkeypatch.setattr(rb, "Pipeline", FakePipeline)
    for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "B200R_BENCH_DEPTH", "B200R_E2E_DEPTH", "B200R_BENCH_FLUSH", "B200R_BENCH_FAKE_SHARD"):
        monkeypatch.delenv(v, raising=False)
    args = types.SimpleNamespace(gpus=1, steps=10, warmup=5, impl="b200", workload="c2", no_cpu_baseline=True)
    bench.run_b200_arm(args, bench.WORKLOADS["c2"])
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    d = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "serial", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "Mrays/s" and d["n_gpus"] == 1 and d["steps"] == 10 and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["config"]["frames_in_flight"] == bench.DEFAULT_DEPTH and "flushed before every frame" in d["config"]["l2"]
    rays = 10 * 1500
    assert d["ms_per_step"] == pytest.approx(0.25 / 10) and d["value"] == pytest.approx(rays / (0.25 / 1000.0) / 1e6)
    alg = 32 * 45000 + 84 * 9000 + 4 * 1920 * 1080
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["algorithmic_bytes_per_launch"] == pytest.approx(alg) and r["kernel_ms"] == pytest.approx(0.6)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and r["achieved"] == pytest.approx(alg / 0.6e-3 / 1e9)
    assert r["concurrency"] == pytest.approx(0.6 / 0.025) and r["alone"]["kernel_ms"] == pytest.approx(0.25)
    assert r["alone"]["frac"] == pytest.approx(d["serial"]["roofline"]["frac"])
    assert d["e2e"]["frames_in_flight"] == bench.DEFAULT_E2E_DEPTH and d["e2e"]["d2h_bytes_per_step"] == 1920 * 1080 * 4
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    # protocol order: a pipeline for `value` (every slot warmed: >= 2 x depth untimed submits, then K), closed, then the e2e pipeline
    pipes = [e for e in log if isinstance(e, tuple) and e[0] == "pipeline"]
    assert pipes == [("pipeline", bench.DEFAULT_DEPTH), ("pipeline", bench.DEFAULT_E2E_DEPTH)]
    first_close = log.index(("pipeline.close", bench.DEFAULT_DEPTH))
    dev = [e for e in log[:first_close] if isinstance(e, tuple) and e[0] == "submit"]
    host = [e for e in log[first_close:] if isinstance(e, tuple) and e[0] == "submit"]
    assert len(dev) == max(5, 2 * bench.DEFAULT_DEPTH) + 10 and not any(e[2] for e in dev)
    assert len(host) == max(5, 2 * bench.DEFAULT_E2E_DEPTH) + 10 and all(e[2] for e in host)
    assert log[-1] == "renderer.close"
    assert "secondary" not in d                                   # quick runs (--no-cpu-baseline) do not append the rasteriser line
    # the full default run: CPU reference leg + the rasteriser line from a process of its own, started after the device was released
    del log[:]
    monkeypatch.setattr(bench, "cpu_reference", lambda wl, target_seconds=15.0, rays_per_frame=None, runner=None:
                        {"value": 60.0, "unit": "Mrays/s", "fps": 27.0, "cores": 16, "kind": "reference", "sample": "stand-in"})
    monkeypatch.setattr(bench, "secondary_line", lambda workload: (log.append(("secondary", workload)), {"metric": "fps", "value": 1900.0})[1])
    args.no_cpu_baseline = False
    bench.run_b200_arm(args, bench.WORKLOADS["c2"])
    d = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "reference" and d["secondary"] == {"metric": "fps", "value": 1900.0}
    assert log.index("renderer.close") < log.index(("secondary", "c4")) and log.count("renderer.close") == 1


def test_b200_arm_sharded_protocol_on_stand_ins(rb, pyport, monkeypatch, capsys):
    """The same dry run as rank 0 of 2 (torch.distributed, NCCL id hand-out and the pipelines are stand-ins): one NCCL id per pipeline,
    8 frames in flight per rank, three pipelines in sequence (serial at depth 1, `value`, e2e), each closed behind a barrier."""
    import types
    import bench
    if not os.path.exists(pyport.model_path("chessboard.tri")):
        pytest.skip("model not staged")
    log = []

    class T:                                                        # a tensor that is a list
        def __init__(self, v): self.v = list(v)
        def clone(self): return T(self.v)
        def tolist(self): return list(self.v)
        def item(self): return self.v[0]
        def __getitem__(self, i): return self.v[i]
        def data_ptr(self): return 4096
        def zero_(self): return self
        def pin_memory(self): return self

    class FakeEvent:
        def __init__(self, enable_timing=False): pass
        def record(self, stream=None): pass
        def elapsed_time(self, other): return 0.5

    fake_torch = types.ModuleType("torch")
    fake_torch.int32, fake_torch.uint8, fake_torch.float64, fake_torch.int64 = "i32", "u8", "f64", "i64"
    fake_torch.zeros = lambda *a, **k: T([0])
    fake_torch.empty = lambda *a, **k: T([0])
    fake_torch.tensor = lambda v, **k: T(v)
    fake_torch.zeros_like = lambda t: T([0] * len(t.v))
    fake_torch.cuda = types.SimpleNamespace(set_device=lambda d: None, Stream=lambda: types.SimpleNamespace(cuda_stream=77),
                                            set_stream=lambda s: None, synchronize=lambda: None, Event=FakeEvent)
    monkeypatch.setitem(sys.modules, "torch", fake_torch)

    def all_reduce(t, op="sum"):
        if op == "sum": t.v = [2 * x for x in t.v]                  # two identical ranks
    def all_gather(out, t):
        for o in out: o.v = list(t.v)
    def broadcast_object_list(lst, src=0): log.append("uid"); assert lst[0] is not None
    fake_dist = types.SimpleNamespace(barrier=lambda: log.append("barrier"), all_reduce=all_reduce, all_gather=all_gather,
                                      broadcast_object_list=broadcast_object_list, ReduceOp=types.SimpleNamespace(MAX="max"),
                                      destroy_process_group=lambda: log.append("destroy_pg"))
    import renderer_b200.dist as rdist
    monkeypatch.setattr(rdist, "init_nccl", lambda local: fake_dist, raising=False)
    monkeypatch.setattr(rb, "dist_unique_id", lambda: b"\0" * 128)
    counters = dict(rays_primary=1000, rays_shadow=500, rays_reflection=0, rays_ao=0, node_tests=40000, leaf_visits=5000, tri_tests=9000,
                    tris_setup=0, spans=0, z_tests=0, z_passes=0)

    class FakeRenderer:
        def __init__(self, device): pass
        def upload(self, scene): pass
        def set_counters(self, on): pass
        def render_device(self, frame, ptr, stream): pass
        def counters(self): return dict(counters)
        def last_launches(self): return 1
        def close(self): log.append("renderer.close")

    class FakePipeline:
        def __init__(self, renderer, width, height, depth=2, rank=0, world=1, unique_id=None, assemble=rb.ASSEMBLE_PUSH):
            assert world == 2 and rank == 0 and unique_id is not None and assemble == rb.ASSEMBLE_PUSH
            self.depth = depth; log.append(("pipeline", depth))
        def submit(self, frame, host=None): assert frame.row_step == 1; log.append(("submit", self.depth, host is not None))
        def drain(self): pass
        def fence(self, stream, pipeline_waits): pass
        def set_l2_flush(self, nbytes, prefetch_scene=True): pass
        def launches(self, reset=False): return 7
        def set_timing(self, enabled): pass
        def kernel_ms(self): return (3.0, 10)
        def close(self):
            assert log[-1] == "barrier", "a pipeline may only be destroyed behind a barrier (peers still write into its buffers)"
            log.append(("pipeline.close", self.depth))
    monkeypatch.setattr(rb, "Renderer", FakeRenderer)
    monkeypatch.setattr(rb, "Pipeline", FakePipeline)
    for v in ("B200R_BENCH_DEPTH", "B200R_E2E_DEPTH", "B200R_BENCH_FLUSH", "B200R_BENCH_FAKE_SHARD", "B200R_ASSEMBLE"):
        monkeypatch.delenv(v, raising=False)
    monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("LOCAL_RANK", "0")
    args = types.SimpleNamespace(gpus=2, steps=10, warmup=5, impl="b200", workload="c2", no_cpu_baseline=False)
    bench.run_b200_arm(args, bench.WORKLOADS["c2"])
    d = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][0])
    D = bench.DEFAULT_DEPTH_SHARDED
    assert d["n_gpus"] == 2 and d["config"]["frames_in_flight"] == D and "row-cyclic sharding over 2 GPUs" in d["config"]["parallelism"]
    assert d["cpu_baseline"] is None and d["e2e"]["frames_in_flight"] == D            # the CPU leg runs at N = 1 only
    assert d["value"] == pytest.approx(2 * 10 * 1500 / (0.5 / 1000.0) / 1e6)          # whole-job rays of both ranks / max-over-ranks time
    assert d["roofline"]["algorithmic_bytes_per_launch"] == pytest.approx(32 * 45000 + 84 * 9000 + 4 * 1920 * 540)   # this rank's rows
    assert "per_rank" in d and len(d["per_rank"]["serial_step_ms"]) == 2
    assert [e for e in log if isinstance(e, tuple) and e[0] == "pipeline"] == [("pipeline", 1), ("pipeline", D), ("pipeline", D)]
    assert log.count("uid") == 3 and log[-2:] == ["destroy_pg", "renderer.close"]
    assert sum(1 for e in log if isinstance(e, tuple) and e[0] == "submit" and e[1] == D and e[2]) == max(5, 2 * D) + 10


def test_secondary_line_condenses_the_rasteriser_run_and_never_raises(monkeypatch):
    """The C4 object a default run appends: the child invocation is `bench.py --workload c4 ... --no-secondary` (no recursion), its JSON
    line (here: the committed one of the final GPU session) is condensed to the headline figures, and any failure becomes
    {"unavailable": ...} instead of an exception."""
    import types
    import bench
    canned = open(os.path.join(ROOT, "profiles", "r03_bench_c4.json")).read()
    calls = []

    def fake_run(cmd, **kw):
        calls.append(cmd)
        return types.SimpleNamespace(returncode=0, stdout="noise\n" + canned + "\n", stderr="")
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    s = bench.secondary_line("c4")
    cmd = calls[0]
    assert cmd[1].endswith("bench.py") and cmd[cmd.index("--workload") + 1] == "c4" and "--no-secondary" in cmd and "--no-cpu-baseline" in cmd
    full = json.loads(canned)
    assert s["metric"] == "fps" and s["value"] == full["value"] and s["workload"].startswith("statue.ply 3840x2160 mode 6")
    assert s["serial"]["fps"] == full["serial"]["fps"] and s["e2e"]["d2h_bytes_per_step"] == 3840 * 2160 * 4 and "clocks" in s
    monkeypatch.setattr(bench.subprocess, "run", lambda cmd, **kw: types.SimpleNamespace(returncode=3, stdout="", stderr="boom"))
    assert "exit 3" in bench.secondary_line("c4")["unavailable"]

    def raising(cmd, **kw):
        raise OSError("no such interpreter")
    monkeypatch.setattr(bench.subprocess, "run", raising)
    assert "OSError" in bench.secondary_line("c4")["unavailable"]
