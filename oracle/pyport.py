"""ctypes access to the CPU restatement (oracle/port -> liboracle_port.so) and to the reference
binaries built by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs; never by renderer_b200.
"""
import ctypes as C
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from renderer_b200 import _abi  # noqa: E402  (POD struct mirrors of include/b200render.h only)

PORT_LIB = os.path.join(HERE, "liboracle_port.so")
REF_BIN = os.path.join(HERE, "_ref", "bin")
MODELS = os.path.join(HERE, "_ref", "models")

PORT_FLAGS = ["-std=c++17", "-O3", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off"]


def build_port(force=False):
    """gcc-compile the C++ restatement (strict IEEE: no fast-math, no FMA contraction)."""
    srcs = sorted(glob.glob(os.path.join(HERE, "port", "*.cpp")))
    deps = srcs + glob.glob(os.path.join(HERE, "port", "*.h")) + [os.path.join(ROOT, "include", "b200render.h")]
    if (not force and os.path.exists(PORT_LIB)
            and os.path.getmtime(PORT_LIB) > max(os.path.getmtime(d) for d in deps)):
        return PORT_LIB
    r = subprocess.run(["g++"] + PORT_FLAGS + ["-o", PORT_LIB] + srcs, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle port build failed:\n" + r.stdout)
    return PORT_LIB


class OScene(C.Structure):
    _fields_ = [("verts", C.POINTER(_abi.Vertex)), ("n_verts", C.c_uint32),
                ("tris", C.POINTER(_abi.Tri)), ("n_tris", C.c_uint32),
                ("nodes", C.POINTER(_abi.BvhNode)), ("n_nodes", C.c_uint32),
                ("tri_idx", C.POINTER(C.c_int32)), ("n_tri_idx", C.c_uint32),
                ("shadowmap", C.c_void_p * 2)]


_port = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_LIB):
            build_port()
        L = C.CDLL(PORT_LIB)
        L.oracle_render.restype = C.c_int
        L.oracle_render.argtypes = [C.POINTER(OScene), C.POINTER(_abi.Frame), C.c_void_p,
                                    C.POINTER(_abi.Counters), C.c_int]
        L.oracle_render_shadowmap.restype = C.c_int
        L.oracle_render_shadowmap.argtypes = [C.POINTER(OScene), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        L.oracle_mlaa.restype = C.c_int
        L.oracle_mlaa.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _port = L
    return _port


def oscene(scene, shadowmaps=()):
    """Wrap a renderer_b200.Scene's host arrays (no copy). Keep `scene` and the maps alive."""
    v, nv, t, nt, n, nn, i, ni = scene.raw()
    s = OScene()
    s.verts, s.n_verts, s.tris, s.n_tris = v, nv, t, nt
    s.nodes, s.n_nodes, s.tri_idx, s.n_tri_idx = n, nn, i, ni
    for k, m in enumerate(shadowmaps):
        s.shadowmap[k] = m.ctypes.data if m is not None else None
    s._keep = (scene, shadowmaps)
    return s


def render(scene, frame, shadowmaps=(), threads=0, counters=False):
    """CPU restatement of one frame -> (rows x width uint32) [and counters dict]."""
    s = oscene(scene, shadowmaps)
    step = frame.row_step or 1
    rows = (frame.height - frame.row_first + step - 1) // step
    out = np.zeros((rows, frame.width), dtype=np.uint32)
    ctr = _abi.Counters()
    rc = port().oracle_render(C.byref(s), C.byref(frame), out.ctypes.data, C.byref(ctr), threads)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    return (out, ctr.as_dict()) if counters else out


def supports(mode, mlaa=False):
    """Which modes the restatement implements (grows as port/*.cpp grows)."""
    L = port()
    L.oracle_supports.restype = C.c_int
    L.oracle_supports.argtypes = [C.c_int, C.c_int]
    return bool(L.oracle_supports(int(mode), 1 if mlaa else 0))


def shadowmaps_for(scene, frame):
    """Shadow maps of the frame's lights, rendered by the restatement of Light::RenderSceneIntoShadowBuffer."""
    import renderer_b200 as rb
    maps = []
    for i in range(frame.n_lights):
        lp = tuple(frame.lights[i].pos)
        w2l = (C.c_float * 9)()
        rb.lib().b200r_light_world_to_light((C.c_float * 3)(*lp), w2l)
        maps.append(render_shadowmap(scene, lp, tuple(w2l)))
    return tuple(maps)


def render_shadowmap(scene, light_pos, world2light):
    s = oscene(scene)
    m = np.empty((1024, 1024), dtype=np.float32)
    rc = port().oracle_render_shadowmap(C.byref(s), (C.c_float * 3)(*light_pos), (C.c_float * 9)(*world2light),
                                        m.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_render_shadowmap failed: {rc}")
    return m


def mlaa(frame_u32):
    fb = np.ascontiguousarray(frame_u32, dtype=np.uint32).copy()
    rc = port().oracle_mlaa(fb.ctypes.data, fb.shape[1], fb.shape[0])
    if rc != 0:
        raise RuntimeError(f"oracle_mlaa failed: {rc}")
    return fb


# ---------------------------------------------------------------- the real reference, built headless

def ref_tag(w, h, no_reflections=False, ao=0, mlaa=False, fast=False, libc_rand=False):
    t = f"{w}x{h}"
    if no_reflections:
        t += "_norefl"
    if ao:
        t += f"_ao{ao}"
    if mlaa:
        t += "_mlaa"
    if fast:
        t += "_fast"
    if libc_rand:
        t += "_libcrand"
    return t


def ref_exe(w, h, **kw):
    return os.path.join(REF_BIN, "renderer_" + ref_tag(w, h, **kw))


def have_ref(w, h, **kw):
    return os.path.exists(ref_exe(w, h, **kw))


def build_ref(w, h, no_reflections=False, ao=0, mlaa=False, fast=False, libc_rand=False):
    """Only possible where /root/reference is mounted (not on the GPU box)."""
    cmd = [sys.executable, os.path.join(HERE, "build_ref.py"), "--w", str(w), "--h", str(h)]
    if no_reflections:
        cmd.append("--no-reflections")
    if ao:
        cmd += ["--ao", str(ao)]
    if mlaa:
        cmd.append("--mlaa")
    if fast:
        cmd.append("--fast")
    if libc_rand:
        cmd.append("--libc-rand")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return ref_exe(w, h, no_reflections=no_reflections, ao=ao, mlaa=mlaa, fast=fast, libc_rand=libc_rand)


def model_path(name):
    return os.path.join(MODELS, name)


def run_ref(model, mode, w, h, frames, two_lights=False, env=None, threads=None, **kw):
    """Run the reference's own `-b` benchmark orbit and return ({frame: array}, stdout).

    The model is used from oracle/_ref/models (writable: the reference drops its .bvh cache there)."""
    exe = ref_exe(w, h, **kw)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    frames = sorted(set(int(f) for f in frames))
    with tempfile.TemporaryDirectory() as td:
        e = dict(os.environ)
        e["ORACLE_DUMP"] = os.path.join(td, "f")
        e["ORACLE_FRAMES"] = ",".join(str(f) for f in frames)
        if threads:
            e["OMP_NUM_THREADS"] = str(threads)
        if env:
            e.update(env)
        cmd = [exe, "-b", "-n", str(frames[-1] + 1), "-m", str(mode % 10)]
        if two_lights:
            cmd.append("-w")
        cmd.append(os.path.basename(model))
        r = subprocess.run(cmd, cwd=os.path.dirname(model), env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, errors="replace")
        if r.returncode != 0:
            raise RuntimeError(f"reference failed ({r.returncode}): {r.stdout[-2000:]}")
        out = {}
        for f in frames:
            out[f] = np.fromfile(os.path.join(td, f"f_{f}.xrgb"), dtype=np.uint32).reshape(h, w)
    return out, r.stdout


def ref_fps(stdout):
    """Parse 'Rendering N frames in S seconds. (F fps)' (reference src/renderer.cc:631-633)."""
    import re
    m = re.search(r"Rendering (\d+) frames in ([0-9.eE+-]+) seconds\. \(([0-9.eE+-]+|inf) fps\)", stdout)
    if not m:
        return None
    n, s = int(m.group(1)), float(m.group(2))
    return n, s, (n / s if s > 0 else float("inf"))
