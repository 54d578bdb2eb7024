#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native renderer hot path.

Metric (BASELINE.json): Mrays/s (and fps) at 1920x1080, at 1/2/4/8 GPUs, next to the reference's own
OpenMP CPU path timed on the same box.  Workload = BASELINE config C2: chessboard.tri, mode 9
(ray tracing: primary rays + Phong + one hard shadow ray per light, reflections off), one light, the
reference's `-b` benchmark orbit (frame k of the orbit is step k).  One "step" = one frame.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c2r|c3|c5]

N > 1 (launched by torchrun, one rank per GPU): the frame is sharded row-cyclically, every rank renders
H/N rows, ONE NCCL all-gather of the packed rows, then a de-interleave kernel (strong scaling).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# One hardware work queue per stream of the frame pipeline (up to 8 render streams + push + consume + the timing stream): the
# CUDA default of 8 connections would let streams share a queue, and a rank's arrival-wait kernel could then hold back the next
# frame's render kernel queued behind it (false dependency; never a deadlock - every dependency points to earlier-submitted work).
# Must be set before CUDA is initialised; measured on one GPU: no effect on the numbers (session r03d).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOADS = {
    # name: model, W, H, mode, flags (1 shadows, 2 reflections, 4 phong normal, 8 AO), ao samples, ref-variant kwargs
    "c2": dict(model="chessboard.tri", W=1920, H=1080, mode=9, flags=1 | 4, ao=0, ref=dict(no_reflections=True),
               desc="chessboard.tri 1920x1080 mode 9: primary rays + Phong + 1 shadow ray, reflections off, 1 light"),
    "c2r": dict(model="chessboard.tri", W=1920, H=1080, mode=9, flags=1 | 2 | 4, ao=0, ref=dict(),
                desc="chessboard.tri 1920x1080 mode 9 reference defaults (reflections on)"),
    "c3": dict(model="dragon_vis.ply", W=1920, H=1080, mode=9, flags=1 | 2 | 4 | 8, ao=16, ref=dict(ao=16),
               desc="dragon_vis.ply 1920x1080 mode 9 + 16-sample AO + reflections"),
    "c5": dict(model="chessboard.tri", W=3840, H=2160, mode=9, flags=1 | 2 | 4 | 8, ao=16, ref=dict(ao=16),
               desc="chessboard.tri 3840x2160 mode 9 + reflections + 16-sample AO"),
    # rasteriser (BASELINE config 4): the metric of this line is fps (there are no rays); flags 0x10 = MLAA post filter
    "c4": dict(model="statue.ply", W=3840, H=2160, mode=6, flags=1 | 2 | 4 | 0x10, ao=0, ref=dict(mlaa=True), raster=True,
               desc="statue.ply 3840x2160 mode 6: per-pixel Phong scan-conversion rasteriser + Z-buffer + MLAA"),
    "c4g": dict(model="statue.ply", W=3840, H=2160, mode=5, flags=1 | 2 | 4 | 0x10, ao=0, ref=dict(mlaa=True), raster=True,
                desc="statue.ply 3840x2160 mode 5: Gouraud scan-conversion rasteriser + Z-buffer + MLAA"),
}


# frames in flight (bench protocol "pipelined" and the e2e leg); B200R_BENCH_DEPTH / B200R_E2E_DEPTH override.
# Measured on one B200 with rt_pool_kernel (profiles/README.md, sessions r03b-e; every slot warmed before the timed region - the
# earlier "6 or 8 in flight lose" was the first-use cudaMalloc of slots 4.. inside it): C2, 4-warp CTAs: 2 in flight 3950 fps,
# 3: 4410, 4-8: 4510-4550. With the frame's rows dealt over N GPUs each rank's launch is short and latency-bound, so more frames
# are needed to fill the GPU: one rank of 8 emulated (every 8th row) 2 in flight 10 100 fps, 4: 13 800, 6: 16 070, 8: 17 320, 12: 17 770.
# e2e (8.3 MB per frame out over PCIe, a slot is busy until its frame has left): 2 slots 3250 fps, 3: 4530, 4: 4770-4980.
DEFAULT_DEPTH = 4
DEFAULT_DEPTH_SHARDED = 8
DEFAULT_E2E_DEPTH = 4
FLUSH_BYTES = 144 << 20      # > 126 MB L2


def algorithmic_bytes(c, W, rows, raster=False):
    """32 B per node popped (inner test or leaf visit), 68 B per triangle tested (4-B index +
    64 B of plane/edge data) + 16 B (centre + twoSided) because culling is on for every ray, + 4 B per pixel.
    Rasteriser (same section): 3 vertices of 28 B + the 144-B triangle per triangle set up, 8 B per z-test,
    4 B per z-pass, and the two clears (frame + depth) of 4 B per pixel each."""
    if raster:
        return (28 * 3 + 144) * c["tris_setup"] + 8 * c["z_tests"] + 4 * c["z_passes"] + 2 * 4 * W * rows
    return 32 * (c["node_tests"] + c["leaf_visits"]) + (68 + 16) * c["tri_tests"] + 4 * W * rows


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    The timed region is short (tens of ms), so NVML is polled from a thread every ~1 ms; `nvidia-smi -lms` (100 ms
    granularity, ~1 s start-up) is only the fallback when NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.thread = None
        self.samples = []           # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = False
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv = nv

            def poll():
                while not self._stop:
                    try:
                        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                             int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.001)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self.nv
            names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
            sm = sorted(s_[0] for s_ in self.samples)
            reasons = sorted({n for _, bits in self.samples for n, b in names if bits & b})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "NVML polled every ~1 ms during the timed region"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# --------------------------------------------------------------------------- CPU reference arm

class RefRunner:
    """The unmodified reference (oracle/_ref/bin/renderer_<cfg>_fast: its own flags -O3 -ffast-math -mrecip
    -fopenmp + SSE paths) run as `renderer -b -n N -m <mode> <model>`; fps from its own printout
    (src/renderer.cc:631-633).  The per-scanline OpenMP fork/join (src/Raytracer.cc:558) does not scale to
    very many threads, so the thread count is calibrated once (all cores, half, 32, 16) and the best is used
    and reported as `cores`."""

    def __init__(self, wl):
        from oracle import pyport
        self.pp = pyport
        self.wl = wl
        self.model = pyport.model_path(wl["model"])
        kw = dict(wl["ref"]); kw["fast"] = True
        self.exe = pyport.ref_exe(wl["W"], wl["H"], **kw)
        self.ok = os.path.exists(self.exe) and os.path.exists(self.model)
        self.threads = os.cpu_count() or 1
        self.fps0 = None

    def run(self, n, threads=None):
        env = dict(os.environ); env["OMP_NUM_THREADS"] = str(threads or self.threads)
        t = time.time()
        r = subprocess.run([self.exe, "-b", "-n", str(n), "-m", str(self.wl["mode"] % 10), os.path.basename(self.model)],
                           cwd=os.path.dirname(self.model), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, errors="replace")
        wall = time.time() - t
        got = self.pp.ref_fps(r.stdout)
        return (got[2] if got and got[1] > 0 else n / max(wall, 1e-3))

    def calibrate(self):
        self.run(1)                               # builds/loads the .bvh cache, warms the page cache
        nproc = os.cpu_count() or 1
        best = (0.0, nproc)
        for th in sorted({nproc, max(1, nproc // 2), min(32, nproc), min(16, nproc)}, reverse=True):
            fps = self.run(3, th)
            if fps > best[0]:
                best = (fps, th)
        self.fps0, self.threads = best
        return best

    def sample(self, seconds):
        n = int(min(400, max(3, (self.fps0 or 3.0) * seconds)))
        return n, self.run(n)


def cpu_reference(wl, target_seconds=15.0, rays_per_frame=None, runner=None):
    """Time the reference's own CPU implementation of the path on this box's host cores -> cpu_baseline object.
    kind "reference" = the unmodified reference binary; falls back to the C++ restatement (kind "port") only
    if that binary was not built."""
    from oracle import pyport
    if rays_per_frame is None:
        rays_per_frame = 0.0
    r = runner or RefRunner(wl)
    if r.ok:
        if r.fps0 is None:
            r.calibrate()
        n, fps = r.sample(target_seconds)
        return {"value": fps if wl.get("raster") else rays_per_frame * fps / 1e6, "unit": "fps" if wl.get("raster") else "Mrays/s",
                "fps": fps, "cores": r.threads,
                "host_cpus": os.cpu_count(), "kind": "reference",
                "sample": f"{n} orbit frames of [{wl['desc']}] by the unmodified reference built with its own flags "
                          f"(-O3 -ffast-math -mrecip -fopenmp, SSE paths); OMP_NUM_THREADS={r.threads} (best of "
                          f"all/half/32/16 on {os.cpu_count()} host CPUs); fps from its own printout"}
    import renderer_b200 as rb
    W, H = wl["W"], wl["H"]
    model = pyport.model_path(wl["model"])
    scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
    cams = rb.Orbit.cameras(range(64))
    t0 = time.time(); n = 0
    while time.time() - t0 < target_seconds and n < 64:
        f = rb.make_frame(wl["mode"], W, H, cams[n], flags=wl["flags"], ao_samples=wl["ao"] or 32, frame_index=n)
        pyport.render(scene, f); n += 1
    fps = n / (time.time() - t0)
    return {"value": fps if wl.get("raster") else rays_per_frame * fps / 1e6, "unit": "fps" if wl.get("raster") else "Mrays/s",
            "fps": fps, "cores": os.cpu_count(), "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{n} orbit frames of [{wl['desc']}] by oracle/port (C++ restatement, OpenMP over rows)"}


def fixture_rays_per_frame(workload, frames):
    """Rays per orbit frame (mean over `frames`) from tests/golden/rays_per_frame.json - counted once by the CPU restatement
    (tests/golden/make_rays_per_frame.py); frames between two stored ones are interpolated. Nothing of the product is loaded."""
    with open(os.path.join(ROOT, "tests", "golden", "rays_per_frame.json")) as f:
        d = {int(k): float(v) for k, v in json.load(f)[workload].items()}
    ks = sorted(d)

    def at(k):
        if k <= ks[0]:
            return d[ks[0]]
        if k >= ks[-1]:
            return d[ks[-1]]
        hi = next(i for i in ks if i >= k)
        lo = max(i for i in ks if i <= k)
        return d[lo] if hi == lo else d[lo] + (d[hi] - d[lo]) * (k - lo) / (hi - lo)
    return sum(at(k) for k in frames) / len(frames)


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    raster = bool(wl.get("raster"))
    # rays per frame: mean over the orbit frames the B200 arm times (warmup .. warmup+steps-1) - for C2 the fixture holds every one of
    # them, so both arms convert fps to Mrays/s with the same count (the restatement's counters equal the device's, tests/)
    K, Wm = max(1, args.steps), args.warmup
    rays = 0 if raster else fixture_rays_per_frame(args.workload, list(range(Wm, Wm + K)))
    # each "step" is a bounded sample: the reference renders a batch of orbit frames; K+W batches in total
    per_step = float(os.environ.get("B200R_REF_STEP_SECONDS", "0")) or max(1.5, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals = []
    runner = RefRunner(wl)
    for i in range(args.warmup + args.steps):
        cb = cpu_reference(wl, target_seconds=per_step, rays_per_frame=rays, runner=runner)
        if i >= args.warmup:
            vals.append(cb)
    fps = sum(v["fps"] for v in vals) / len(vals)
    unit = "fps" if raster else "Mrays/s"
    val = fps if raster else rays * fps / 1e6
    per_step = sorted(v["fps"] for v in vals)
    cb = dict(vals[-1]); cb["value"] = val; cb["fps"] = fps
    cb["fps_median"] = per_step[len(per_step) // 2]; cb["fps_best"] = per_step[-1]; cb["fps_worst"] = per_step[0]
    line = {"impl": "reference", "metric": unit, "value": val, "unit": unit, "fps": fps, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / fps if fps else None,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # (the same keys as the B200 arm's `config`, so the two lines can be laid side by side)
            "config": {"workload": wl["desc"], "camera": "reference -b orbit, one new frame per step", "rays_per_frame": rays,
                       "raster_per_frame": None, "l2": "n/a: reference CPU arm, no GPU involved", "frames_in_flight": 1,
                       "host_enqueue_ms_per_step": None, "parallelism": f"{cb['cores']} host threads (the reference's OpenMP loop)",
                       "timing": "fps from the reference's own printout (src/renderer.cc:631-633) over a batch of orbit frames per step"},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- B200 arm

def secondary_line(workload, steps=20, timeout=150):
    """BASELINE config 4 (the scan-conversion rasteriser + Z-buffer + MLAA at 3840x2160) beside the headline ray-tracing line, so that a
    driver-run rasteriser number exists: the very same protocol (`python bench.py --workload c4`), run in a process of its own after
    this one has released the device, condensed to its headline figures. Never fails the main line: anything unexpected is reported
    as {"unavailable": why}."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", workload, "--steps", str(steps), "--warmup", "5",
                            "--no-cpu-baseline", "--no-secondary"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not rows:
            return {"unavailable": ("exit %d: " % r.returncode) + r.stderr.strip()[-300:]}
        d = json.loads(rows[-1])
        sr = d["serial"]["roofline"]
        return {"workload": d["config"]["workload"], "metric": d["metric"], "value": d["value"], "unit": d["unit"], "fps": d["fps"],
                "steps": d["steps"], "warmup": d["warmup"], "ms_per_step": d["ms_per_step"], "frames_in_flight": d["config"]["frames_in_flight"],
                "l2": d["config"]["l2"], "gpu_launches": d["gpu_launches"],
                "serial": {"fps": d["serial"]["fps"], "ms_per_step": d["serial"]["ms_per_step"], "kernel_ms": sr["kernel_ms"],
                           "algorithmic_bytes_per_launch": sr["algorithmic_bytes_per_launch"], "frac": sr["frac"]},
                "e2e": {k: d["e2e"][k] for k in ("value", "unit", "fps", "h2d_bytes_per_step", "d2h_bytes_per_step", "frames_in_flight")},
                "clocks": d["clocks"], "how": "python bench.py --workload %s --steps %d --warmup 5 --no-cpu-baseline, own process" % (workload, steps)}
    except Exception as e:           # noqa: BLE001 - the headline line must be printed whatever happens here
        return {"unavailable": ("%s: %s" % (type(e).__name__, e))[:300]}


def run_b200_arm(args, wl):
    import ctypes as C
    import numpy as np
    import torch
    import renderer_b200 as rb
    from oracle import pyport   # only for model staging paths and the cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    P = world
    torch.cuda.set_device(local)
    dist = None
    if P > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        from renderer_b200.dist import init_nccl
        dist = init_nccl(local)

    W, H = wl["W"], wl["H"]
    raster = bool(wl.get("raster"))
    model = pyport.model_path(wl["model"])
    scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
    gpu = rb.Renderer(local)
    gpu.upload(scene)

    K, Wm = args.steps, args.warmup
    cams = rb.Orbit.cameras(range(K + Wm))
    rows_per = (H + P - 1) // P
    assemble = {"nccl": rb.ASSEMBLE_NCCL, "push": rb.ASSEMBLE_PUSH}[os.environ.get("B200R_ASSEMBLE", "push")]

    def frame_for(step, sharded=False):
        return rb.make_frame(wl["mode"], W, H, cams[step], flags=wl["flags"], ao_samples=wl["ao"] or 32,
                             frame_index=step, row_first=rank if sharded else 0, row_step=P if sharded else 1)

    def job_id():
        """One NCCL unique id per pipeline, created on rank 0 and handed out over torch.distributed (plumbing)."""
        if P == 1:
            return None
        uid = [rb.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        return uid[0]

    # Everything timed is ordered against ONE explicit stream: the L2 flush, the events, the renderer's kernels.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    mine = torch.zeros((rows_per if P > 1 else H, W), dtype=torch.int32, device="cuda")      # this rank's rows (device-resident output)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier():
        if P > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- counters of every timed frame (untimed counting pass; counting runs the reference's own traversal order, no pruning)
    gpu.set_counters(True)
    per_frame = []
    for s in range(Wm, Wm + K):
        gpu.render_device(frame_for(s, sharded=P > 1), mine.data_ptr(), None)
        per_frame.append(gpu.counters())
    gpu.set_counters(False)
    keys = ("rays_primary", "rays_shadow", "rays_reflection", "rays_ao", "node_tests", "leaf_visits", "tri_tests",
            "tris_setup", "spans", "z_tests", "z_passes")
    tot = {k: sum(c[k] for c in per_frame) for k in keys}
    if P > 1:
        t = torch.tensor([tot[k] for k in keys], dtype=torch.int64, device="cuda")
        me = t.clone()
        dist.all_reduce(t)
        tot_all = {k: int(v) for k, v in zip(keys, t.tolist())}
        tot_mine = {k: int(v) for k, v in zip(keys, me.tolist())}
    else:
        tot_all = tot_mine = tot
    rays_total = tot_all["rays_primary"] + tot_all["rays_shadow"] + tot_all["rays_reflection"] + tot_all["rays_ao"]
    unit = "fps" if raster else "Mrays/s"
    peak, peak_src = measured_peaks()
    alg_bytes_mine = algorithmic_bytes(tot_mine, W, (rows_per if P > 1 else H) * K, raster)

    def roofline_of(kernel_ms_avg, note):
        achieved = (alg_bytes_mine / K) / (kernel_ms_avg / 1000.0) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": ("rasteriser step = clears + ras_setup + ras_depth + ras_resolve (+ the MLAA kernels)" if raster else
                           "ray-tracing step of this rank = frame clear + rt_pool_kernel (dominant; in the C2 configuration it shades and "
                           "casts the shadow rays itself) [+ rt_shade_kernel for AO / reflections]"),
                "kernel_ms": kernel_ms_avg, "algorithmic_bytes_per_launch": alg_bytes_mine / K,
                "peak_source": peak_src + " (of measured)", "how": note}

    # ---- protocol 1, `serial`: one frame at a time. Per step: L2 flush (outside the events), event, this rank's kernels, event,
    # [N > 1: assembly on every rank], event. This is the run the stand-alone kernel time comes from.
    pipe1 = rb.Pipeline(gpu, W, H, depth=1, rank=rank, world=P, unique_id=job_id(), assemble=assemble) if P > 1 else None
    for s in range(Wm):
        gpu.render_device(frame_for(s, sharded=P > 1), mine.data_ptr(), sptr)
        if pipe1:
            pipe1.submit(frame_for(s))
    if pipe1:
        pipe1.drain()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
    serial_launches = 0
    for i in range(K):
        flush.zero_()                       # evict the 126 MB L2 between timed iterations (not timed)
        ev[i][0].record()
        if P == 1:
            gpu.render_device(frame_for(Wm + i), mine.data_ptr(), sptr)
            serial_launches += gpu.last_launches()
            ev[i][1].record()
        else:
            # the whole step through the pipeline (depth 1): render my rows, assemble on every rank
            ev[i][1].record()
            pipe1.fence(sptr, True)
            pipe1.submit(frame_for(Wm + i))
            pipe1.fence(sptr, False)
        ev[i][2].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    serial_total_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
    if P == 1:
        kern_alone_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    else:
        serial_launches = pipe1.launches(reset=True)
        # this rank's kernels alone (no assembly), same flush protocol: the stand-alone kernel time of a rank's shard
        kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(2)) for _ in range(K)]
        for i in range(K):
            flush.zero_()
            kev[i][0].record()
            gpu.render_device(frame_for(Wm + i, sharded=True), mine.data_ptr(), sptr)
            kev[i][1].record()
        barrier()
        kern_alone_ms = sum(e[0].elapsed_time(e[1]) for e in kev) / K
    per_rank = None
    if P > 1:
        me = torch.tensor([serial_total_ms / K, kern_alone_ms], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(me) for _ in range(P)]
        dist.all_gather(allr, me)
        per_rank = {"serial_step_ms": [float(a[0]) for a in allr], "render_kernels_alone_ms": [float(a[1]) for a in allr]}
        t = torch.tensor([serial_total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        serial_total_ms = float(t.item())
        barrier()
        pipe1.close()
    serial_ms_per_step = serial_total_ms / K

    # ---- protocol 2, the headline `value`: the same K steps with frames IN FLIGHT through b200r_pipeline (the product's mode for
    # independent frames - the reference's own -b loop): frame i+1's kernels start while frame i's tail is still running, and (N > 1)
    # frame i is assembled while the next frames render. The L2 flush stays: one > L2 write enqueued on the frame's stream before every
    # frame, INSIDE the timed region, followed by a bulk L2 prefetch of the scene. One start event, one end event after every stream
    # of the pipeline has joined the timing stream; max over ranks.
    depth = int(os.environ.get("B200R_BENCH_DEPTH", "0")) or (DEFAULT_DEPTH if P == 1 else DEFAULT_DEPTH_SHARDED)
    do_flush = os.environ.get("B200R_BENCH_FLUSH", "1") != "0"
    pipe_ms_per_step = None
    enqueue_ms_per_step = None
    if True:
        pipe = rb.Pipeline(gpu, W, H, depth=depth, rank=rank, world=P, unique_id=job_id(), assemble=assemble)
        if do_flush:
            # (a bulk L2 prefetch of the scene behind the flush - b200r_pipeline_set_prefetch - was measured and LOSES: one rank of 8
            # emulated on one GPU 10 080 fps with it, 11 220 without; B200R_BENCH_PREFETCH=1 switches it on)
            pipe.set_l2_flush(FLUSH_BYTES, prefetch_scene=os.environ.get("B200R_BENCH_PREFETCH", "0") == "1")
        frames = [frame_for(s_) for s_ in range(Wm + K)]         # frame state prepared outside the timed region (12 floats each)
        fake = int(os.environ.get("B200R_BENCH_FAKE_SHARD", "0"))   # developer experiment: one GPU renders rows 0, P, 2P.. only (a rank's load)
        if fake > 1 and P == 1:
            for f_ in frames:
                f_.row_first, f_.row_step = 0, fake
        for s_ in range(max(Wm, 2 * depth)):          # every slot allocates its scratch buffers on first use: warm all of them
            pipe.submit(frames[s_ % Wm])
        pipe.drain()
        barrier()
        pipe.launches(reset=True)
        pipe.set_timing(True)
        sampler2 = ClockSampler(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        e0.record(stream)
        pipe.fence(sptr, True)
        for i in range(K):
            pipe.submit(frames[Wm + i])
        enqueue_ms_per_step = (time.perf_counter() - t_wall0) * 1000.0 / K     # host time to enqueue one frame (all its stages)
        pipe.fence(sptr, False)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t_wall0
        clocks2 = sampler2.stop() if sampler2 else None
        pipe_total_ms = e0.elapsed_time(e1)
        if P > 1:
            t = torch.tensor([pipe_total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pipe_total_ms = float(t.item())
        pipe_ms_per_step = pipe_total_ms / K
        launches = pipe.launches()
        ksum, kn = pipe.kernel_ms()
        pipe.set_timing(False)
        kern_inflight_ms = ksum / max(kn, 1)
        if clocks2 and clocks2.get("sm_mhz"):
            clocks = clocks2
        # ---- end to end through the same public call with HOST buffers: every assembled frame is copied to page-locked host memory
        # of rank 0 (N = 1: of the one rank) inside the timed region; no L2 flush here (the user-facing call has none)
        # (its own pipeline: a slot is busy until its frame has left over PCIe - 8.3 MB, ~0.17 ms at 1080p - so the copy-out wants one
        # more slot than the device-only protocol)
        barrier()
        pipe.close()
        e2e_depth = int(os.environ.get("B200R_E2E_DEPTH", "0")) or (DEFAULT_E2E_DEPTH if P == 1 else depth)
        pipe = rb.Pipeline(gpu, W, H, depth=e2e_depth, rank=rank, world=P, unique_id=job_id(), assemble=assemble)
        hosts = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(e2e_depth)] if rank == 0 else None
        for s_ in range(max(Wm, 2 * e2e_depth)):
            pipe.submit(frames[s_ % Wm], hosts[s_ % e2e_depth].data_ptr() if hosts else None)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            pipe.submit(frames[Wm + i], hosts[i % e2e_depth].data_ptr() if hosts else None)
        pipe.drain()                     # every frame of the timed region is complete in rank 0's host memory
        barrier()
        e2e_s = time.perf_counter() - t0
        if P > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        barrier()
        pipe.close()
        e2e_note = (f"b200r_pipeline_submit with a page-locked host frame per step: frame state in, the assembled XRGB frame out to rank 0's host "
                    f"memory; {e2e_depth} frames in flight per rank; every frame is complete in host memory before the clock stops (b200r_pipeline_drain "
                    "+ barrier); wall clock, max over ranks")
    ms_per_step = pipe_ms_per_step if pipe_ms_per_step is not None else serial_ms_per_step
    fps = 1000.0 / ms_per_step
    value = fps if raster else rays_total / (ms_per_step * K / 1000.0) / 1e6
    e2e_value = K / e2e_s if raster else rays_total / e2e_s / 1e6

    if rank == 0:
        cpu = None
        roof_main = None
        if pipe_ms_per_step is not None:
            roof_main = roofline_of(kern_inflight_ms, f"same run as `value`: CUDA events on each frame's own stream around this rank's kernels of that frame, "
                                                      f"mean over the K timed frames. Up to {depth} frames are in flight, so a launch shares the SMs with its "
                                                      "neighbours: kernel_ms is a RESIDENCE time and exceeds ms_per_step by about `concurrency`; `alone` "
                                                      "(= `serial.roofline`) is the same launch timed by itself in this very invocation")
            roof_main["concurrency"] = kern_inflight_ms / pipe_ms_per_step          # launches of this rank resident at the same time, on average
            a_alone = (alg_bytes_mine / K) / (kern_alone_ms / 1000.0) / 1e9
            roof_main["alone"] = {"kernel_ms": kern_alone_ms, "achieved": a_alone, "frac": a_alone / peak, "unit": "GB/s"}
        if P == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference(wl, target_seconds=15.0, rays_per_frame=rays_total / K)
        in_flight = depth if pipe_ms_per_step is not None else 1
        line = {
            "metric": unit, "value": value, "unit": unit, "fps": fps, "n_gpus": P, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "camera": "reference -b orbit, one new frame per step",
                       "rays_per_frame": rays_total / K,
                       "raster_per_frame": ({k: tot_all[k] / K for k in ("tris_setup", "spans", "z_tests", "z_passes")} if raster else None),
                       "l2": ("flushed before every frame: a 144 MiB write (> 126 MB L2; 16-byte stores from two 128-thread CTAs per SM) enqueued on the frame's stream, INSIDE the timed region"
                              if pipe_ms_per_step is not None and do_flush else
                              ("NOT flushed (B200R_BENCH_FLUSH=0: experiment, not a bench value)" if pipe_ms_per_step is not None else
                               "flushed between timed steps (144 MiB write, outside the events)")),
                       "frames_in_flight": in_flight,
                       **({"EXPERIMENT_fake_shard": int(os.environ["B200R_BENCH_FAKE_SHARD"])} if os.environ.get("B200R_BENCH_FAKE_SHARD") else {}),
                       "host_enqueue_ms_per_step": enqueue_ms_per_step,
                       "parallelism": "1 GPU" if P == 1 else
                                      (f"row-cyclic sharding over {P} GPUs; rows assembled on every rank by " +
                                       ("peer stores over NVLink + one arrival flag per frame (b200r_pipeline, B200R_ASSEMBLE_PUSH)"
                                        if assemble == rb.ASSEMBLE_PUSH else "ONE ncclAllGather + de-interleave per frame (B200R_ASSEMBLE_NCCL)")),
                       "timing": ("exactly K frames between ONE start event (before the first flush) and ONE end event recorded after every "
                                  "stream of the pipeline has joined the timing stream; max over ranks. Frames are independent (the reference's "
                                  "-b orbit), so up to frames_in_flight of them overlap per rank; `serial` is the same K frames one at a time")
                                 if pipe_ms_per_step is not None else
                                 "CUDA events on the launching stream around every step, max over ranks"},
            # one protocol per object: `roofline` belongs to the run `value` comes from (per-launch durations measured in that very run);
            # `serial` is a first-class object with its own ms_per_step, kernel time and roofline.
            "roofline": roof_main if roof_main is not None else
                        roofline_of(kern_alone_ms, "same run as `value`: CUDA events around this rank's kernels of every step"),
            "serial": {"ms_per_step": serial_ms_per_step, "fps": 1000.0 / serial_ms_per_step,
                       "value": (1000.0 / serial_ms_per_step) if raster else rays_total / (serial_ms_per_step * K / 1000.0) / 1e6,
                       "unit": unit, "gpu_launches": serial_launches, "frames_in_flight": 1,
                       "roofline": roofline_of(kern_alone_ms, "one frame at a time, L2 flushed (144 MiB write) before every step outside the events; CUDA "
                                                              "events around this rank's kernels of every step: a launch timed alone"),
                       "note": "one frame at a time on one stream; kernel_ms <= ms_per_step holds here by construction"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": unit, "fps": K / e2e_s,
                    "h2d_bytes_per_step": C.sizeof(rb.Frame), "d2h_bytes_per_step": W * H * 4,
                    "frames_in_flight": e2e_depth, "note": e2e_note,
                    },
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall,
        }
        if per_rank:
            line["per_rank"] = per_rank
        if P == 1 and args.workload == "c2" and not args.no_cpu_baseline and not getattr(args, "no_secondary", False):     # the full default run only
            gpu.close()                  # this process is done with the device: the rasteriser line is measured by a process of its own
            gpu = None
            line["secondary"] = secondary_line("c4")
        print(json.dumps(line))
    if P > 1:
        dist.barrier()
        dist.destroy_process_group()
    if gpu is not None:
        gpu.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the rasteriser (C4) line a default C2 run on one GPU appends")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_b200_arm(args, wl)


if __name__ == "__main__":
    main()
