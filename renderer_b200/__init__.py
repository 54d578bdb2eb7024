"""renderer_b200 — thin Python binding over the C-ABI of libb200render.so (include/b200render.h).

The product is the shared library (C++ host plumbing + hand-written sm_100a CUDA kernels); this
module only loads it with ctypes so that tests and bench.py can drive it. Names follow the
reference's domain (Scene / Camera / Light / Screen and the Scene::render* entry points of
reference src/Scene.h:76-85). There is no Python or CPU rendering fallback: if the library is
missing, importing works but every use raises; if there is no B200, Renderer() raises.
"""
import contextlib
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (F_AO, F_DEFAULT, F_MLAA, F_PHONG_NORMAL, F_REFLECTIONS, F_SHADOWS,  # noqa: F401
                   MODE_AMBIENT, MODE_GOURAUD, MODE_LINES, MODE_PHONG, MODE_PHONG_SHADOWMAPS,
                   MODE_PHONG_SOFTSHADOWMAPS, MODE_POINTS, MODE_POINTS_TRI, MODE_RAYTRACE,
                   MODE_RAYTRACE_AA, MAX_FRAMES_IN_FLIGHT, Counters, Frame)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200R_LIB") or os.path.join(_PKG, "libb200render.so")   # (B200R_LIB: developer builds)
_lib = None


class RendererError(RuntimeError):
    """Raised where the reference would THROW(std::string) (src/Exceptions.h:28-33) or exit()."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RendererError(
                f"{LIB_PATH} is missing: run `python -m renderer_b200.build` (or __graft_entry__.build()). "
                "There is no fallback renderer.")
        _lib = _abi.bind(C.CDLL(LIB_PATH))
    return _lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _f9(v):
    return (C.c_float * 9)(*[float(x) for x in v])


def _check(rc, ctx=None):
    if rc != 0:
        msg = lib().b200r_last_error(ctx)
        raise RendererError(f"[{rc}] {msg.decode() if msg else 'unknown error'}")


class Scene:
    """Host-side scene = reference `Scene` after load(): vertices, triangles, flattened BVH."""

    def __init__(self, filename=None):
        self._h = C.c_void_p()
        self.filename = None
        if filename is not None:
            self.load(filename)

    def load(self, filename):
        """Scene::load (reference src/Loader.cc:85)."""
        if self._h:
            lib().b200r_scene_free(self._h)
            self._h = C.c_void_p()
        _check(lib().b200r_scene_load(os.fsencode(filename), C.byref(self._h)))
        self.filename = filename
        return self

    def UpdateBoundingVolumeHierarchy(self, cache_path=None, forceRecalc=False):
        """Scene::UpdateBoundingVolumeHierarchy (reference src/Raytracer.cc:720). cache_path is the
        `<model>.bvh` file (same format as the reference's); None = build in memory only."""
        _check(lib().b200r_scene_build_bvh(self._h, os.fsencode(cache_path) if cache_path else None,
                                            1 if forceRecalc else 0))
        return self

    def UpdateBoundingVolumeHierarchyOnDevice(self, renderer, cache_path=None, forceRecalc=False):
        """The same, with CreateBVH/CreateCFBVH run as CUDA kernels (b200r_build_bvh): same tree, same cache file."""
        _check(lib().b200r_scene_build_bvh_device(self._h, renderer._ctx, os.fsencode(cache_path) if cache_path else None,
                                                   1 if forceRecalc else 0), renderer._ctx)
        return self

    def _arr(self, fn, ctype):
        n = C.c_uint32()
        p = fn(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint8), 0
        buf = C.cast(p, C.POINTER(C.c_uint8 * (n.value * C.sizeof(ctype)))).contents
        return np.frombuffer(buf, dtype=np.uint8), n.value

    @property
    def n_vertices(self):
        n = C.c_uint32(); lib().b200r_scene_vertices(self._h, C.byref(n)); return n.value

    @property
    def n_triangles(self):
        n = C.c_uint32(); lib().b200r_scene_tris(self._h, C.byref(n)); return n.value

    @property
    def n_nodes(self):
        n = C.c_uint32(); lib().b200r_scene_nodes(self._h, C.byref(n)); return n.value

    @property
    def bvh_depth(self):
        return lib().b200r_scene_bvh_depth(self._h)

    def raw(self):
        """(verts_ptr, nv, tris_ptr, nt, nodes_ptr, nn, triidx_ptr, ni) for the oracle / tests."""
        L = lib()
        nv, nt, nn, ni = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        v = L.b200r_scene_vertices(self._h, C.byref(nv)); t = L.b200r_scene_tris(self._h, C.byref(nt))
        n = L.b200r_scene_nodes(self._h, C.byref(nn)); i = L.b200r_scene_tri_idx(self._h, C.byref(ni))
        return v, nv.value, t, nt.value, n, nn.value, i, ni.value

    def _built_bvh_bytes(self, call):
        v, nv, t, nt, _, _, _, _ = self.raw()
        nodes = (_abi.BvhNode * (2 * nt + 1))()
        idx = (C.c_int32 * nt)()
        nn, depth = C.c_uint32(), C.c_int32()
        call(v, nv, t, nt, nodes, 2 * nt + 1, idx, C.byref(nn), C.byref(depth))
        head = np.array([nn.value, nt], dtype=np.uint32).tobytes()
        return head + bytes(memoryview(nodes))[:nn.value * C.sizeof(_abi.BvhNode)] + bytes(memoryview(idx)), depth.value

    def bvh_bytes_from_steps_on_host(self):
        """(.bvh cache bytes, depth) produced by the device build's step functions run in plain loops on the host (test hook)."""
        return self._built_bvh_bytes(lambda *a: _check(lib().b200r_selftest_bvh_steps_host(*a)))

    def bvh_bytes_from_device_build(self, renderer):
        """(.bvh cache bytes, depth) produced by b200r_build_bvh: the SAH build as CUDA kernels."""
        return self._built_bvh_bytes(lambda *a: _check(lib().b200r_build_bvh(renderer._ctx, *a), renderer._ctx))

    def bvh_bytes(self):
        """The flattened BVH in the reference's .bvh cache layout (Raytracer.cc:747-753)."""
        nodes, nn = self._arr(lib().b200r_scene_nodes, _abi.BvhNode)
        idx, ni = self._arr(lib().b200r_scene_tri_idx, C.c_int32)
        return np.array([nn, ni], dtype=np.uint32).tobytes() + nodes.tobytes() + idx.tobytes()

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.b200r_scene_free(self._h)
        except Exception:
            pass


class Camera:
    """reference `Camera` (src/Camera.h:26-58): eye position + row matrix {up, right, forward}."""

    def __init__(self, eye=(4.8, 0.0, 0.0), lookat=(0.0, 0.0, 0.0)):
        self.eye = (C.c_float * 3)()
        self.mv = (C.c_float * 9)()
        self.set(eye, lookat)

    def set(self, eye, lookat):
        e = _f3(eye)
        lib().b200r_camera_look_at(e, _f3(lookat), self.mv)
        C.memmove(self.eye, e, 12)
        return self


class Orbit:
    """The `-b` benchmark orbit of main() (reference src/renderer.cc:485-496): frame k of the orbit
    is the camera after k+1 calls of step()."""

    def __init__(self):
        self._o = _abi.Orbit()
        lib().b200r_orbit_init(C.byref(self._o))

    def step(self):
        cam = Camera.__new__(Camera)
        cam.eye = (C.c_float * 3)(); cam.mv = (C.c_float * 9)()
        lib().b200r_orbit_step(C.byref(self._o), cam.eye, cam.mv)
        return cam

    @staticmethod
    def cameras(frames):
        """{k: Camera} for the requested orbit frame numbers."""
        frames = sorted(set(int(k) for k in frames))
        o, out = Orbit(), {}
        for k in range(frames[-1] + 1):
            cam = o.step()
            if k in frames:
                out[k] = cam
        return out


def make_frame(mode, width, height, camera, n_lights=1, flags=F_DEFAULT, ao_samples=32, max_depth=3,
               frame_index=0, row_first=0, row_step=1):
    """One iteration of main()'s loop for `mode`: camera + default lights + per-mode light matrices."""
    f = Frame()
    lib().b200r_frame_defaults(C.byref(f), mode, width, height, camera.eye, camera.mv, n_lights)
    f.flags = flags; f.ao_samples = ao_samples; f.max_depth = max_depth; f.frame_index = frame_index
    f.row_first = row_first; f.row_step = row_step
    return f


class Renderer:
    """Device context: replaces `Screen` + the Scene::render* calls (reference src/Scene.h:76-85)."""

    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        _check(lib().b200r_init(int(device), C.byref(self._ctx)))
        self.scene = None

    def close(self):
        if self._ctx:
            lib().b200r_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, scene):
        _check(lib().b200r_upload_scene_handle(self._ctx, scene._h), self._ctx)
        self.scene = scene
        return self

    def upload_shadowmap(self, light, shadowmap):
        m = np.ascontiguousarray(shadowmap, dtype=np.float32)
        assert m.shape == (1024, 1024)
        _check(lib().b200r_upload_shadowmap(self._ctx, light, m.ctypes.data), self._ctx)

    def render_shadowmap(self, light, light_pos):
        w2l = (C.c_float * 9)()
        lp = _f3(light_pos)
        lib().b200r_light_world_to_light(lp, w2l)
        _check(lib().b200r_render_shadowmap(self._ctx, light, lp, w2l), self._ctx)

    def download_shadowmap(self, light):
        m = np.empty((1024, 1024), dtype=np.float32)
        _check(lib().b200r_download_shadowmap(self._ctx, light, m.ctypes.data), self._ctx)
        return m

    def selftest_division(self, samples=1 << 32, seed=1):
        """(#mismatches, first_bad) of the shared-reciprocal divide vs the IEEE divide (must be 0)."""
        m = C.c_uint64()
        fb = (C.c_float * 4)()
        _check(lib().b200r_selftest_division(self._ctx, samples, seed, C.byref(m), fb), self._ctx)
        return m.value, tuple(fb)

    def tile_profile(self, frame):
        """Developer tool: render `frame` with per-tile timestamps -> (n_tiles, 2) uint64 array of start/end ns."""
        _check(lib().b200r_set_tile_profile(self._ctx, 1), self._ctx)
        try:
            self.render(frame)
            n = C.c_uint32()
            _check(lib().b200r_get_tile_profile(self._ctx, None, 0, C.byref(n)), self._ctx)
            out = np.zeros((n.value, 2), dtype=np.uint64)
            _check(lib().b200r_get_tile_profile(self._ctx, out.ctypes.data, n.value, C.byref(n)), self._ctx)
        finally:
            _check(lib().b200r_set_tile_profile(self._ctx, 0), self._ctx)
        return out

    def set_switch(self, name, value=1):
        """Developer switch (b200r_set_switch): selects a cross-check variant of a kernel; results never change."""
        _check(lib().b200r_set_switch(self._ctx, name.encode(), int(value)), self._ctx)

    @contextlib.contextmanager
    def switch(self, name, value=1, restore=0):
        """`with gpu.switch("no_prune"): ...` - the switch is set inside the block and put back to `restore` after it."""
        self.set_switch(name, value)
        try:
            yield self
        finally:
            self.set_switch(name, restore)

    def set_counters(self, enabled):
        _check(lib().b200r_set_counters(self._ctx, 1 if enabled else 0), self._ctx)

    def counters(self):
        c = Counters()
        _check(lib().b200r_get_counters(self._ctx, C.byref(c)), self._ctx)
        return c.as_dict()

    def last_kernel_ms(self):
        t, d = C.c_float(), C.c_float()
        _check(lib().b200r_last_kernel_ms(self._ctx, C.byref(t), C.byref(d)), self._ctx)
        return t.value, d.value

    def last_launches(self):
        n = C.c_uint32()
        _check(lib().b200r_last_launches(self._ctx, C.byref(n)), self._ctx)
        return n.value

    @staticmethod
    def rows_of(frame):
        step = frame.row_step or 1
        return (frame.height - frame.row_first + step - 1) // step

    def render(self, frame, out=None):
        """b200r_render: host buffer out (rows x width uint32 0x00RRGGBB), D2H copy included."""
        rows = self.rows_of(frame)
        if out is None:
            out = np.empty((rows, frame.width), dtype=np.uint32)
        _check(lib().b200r_render(self._ctx, C.byref(frame), out.ctypes.data), self._ctx)
        return out

    def render_async(self, frame, out):
        """b200r_render_async: enqueue the frame; `out` (rows x width uint32, ideally page-locked) is complete after
        wait() or after the (pipeline depth + 1)-th following render_async()."""
        pending = getattr(self, "_pending", []) + [out]
        _check(lib().b200r_render_async(self._ctx, C.byref(frame), out.ctypes.data), self._ctx)
        # the call above retired the frame submitted depth+1 calls ago (its staging copy may land in that buffer during the call):
        # only now may older buffers be let go; the last depth+1 are still in flight
        self._pending = pending[-(getattr(self, "_depth", 2) + 1):]
        return out

    def wait(self):
        _check(lib().b200r_wait(self._ctx), self._ctx)
        self._pending = []

    def set_pipeline_depth(self, depth):
        """b200r_set_pipeline_depth: ray-traced frames of render_async that render concurrently (1..MAX_FRAMES_IN_FLIGHT)."""
        _check(lib().b200r_set_pipeline_depth(self._ctx, int(depth)), self._ctx)
        self._depth = int(depth)
        self._pending = []

    def render_device_slot(self, frame, dev_ptr, stream, slot):
        """b200r_render_device_slot: enqueue a ray-traced frame on `stream` (a cudaStream_t handle, not 0) using scratch set
        `slot`; frames with different slots may be in flight together, nothing is synchronised."""
        _check(lib().b200r_render_device_slot(self._ctx, C.byref(frame), C.c_void_p(dev_ptr), C.c_void_p(stream), int(slot)),
               self._ctx)

    def render_device(self, frame, dev_ptr, stream=None):
        _check(lib().b200r_render_device(self._ctx, C.byref(frame), C.c_void_p(dev_ptr),
                                          C.c_void_p(stream) if stream else None), self._ctx)

    def mlaa_device(self, dev_ptr, width, height, stream=None):
        _check(lib().b200r_mlaa_device(self._ctx, C.c_void_p(dev_ptr), width, height,
                                        C.c_void_p(stream) if stream else None), self._ctx)

    def deinterleave_device(self, gathered_ptr, frame_ptr, width, height, n_shards, stream=None):
        _check(lib().b200r_deinterleave_device(self._ctx, C.c_void_p(gathered_ptr), C.c_void_p(frame_ptr), width,
                                                height, n_shards, C.c_void_p(stream) if stream else None), self._ctx)

    # ---- the reference's entry points, by name (src/Scene.h:76-85) ----
    def _mode(self, mode, camera, width, height, **kw):
        return self.render(make_frame(mode, width, height, camera, **kw))

    def renderPoints(self, camera, width, height, asTriangles=True, **kw):
        return self._mode(MODE_POINTS_TRI if asTriangles else MODE_POINTS, camera, width, height, **kw)

    def renderWireframe(self, camera, width, height, **kw):
        return self._mode(MODE_LINES, camera, width, height, **kw)

    def renderAmbient(self, camera, width, height, **kw):
        return self._mode(MODE_AMBIENT, camera, width, height, **kw)

    def renderGouraud(self, camera, width, height, **kw):
        return self._mode(MODE_GOURAUD, camera, width, height, **kw)

    def renderPhong(self, camera, width, height, **kw):
        return self._mode(MODE_PHONG, camera, width, height, **kw)

    def renderPhongAndShadowed(self, camera, width, height, **kw):
        return self._mode(MODE_PHONG_SHADOWMAPS, camera, width, height, **kw)

    def renderPhongAndSoftShadowed(self, camera, width, height, **kw):
        return self._mode(MODE_PHONG_SOFTSHADOWMAPS, camera, width, height, **kw)

    def renderRaytracer(self, camera, width, height, antiAlias=False, **kw):
        return self._mode(MODE_RAYTRACE_AA if antiAlias else MODE_RAYTRACE, camera, width, height, **kw)


ASSEMBLE_NCCL, ASSEMBLE_PUSH = 0, 1


def dist_unique_id():
    """b200r_dist_unique_id: 128 bytes naming a multi-GPU job; create on rank 0 and hand to every rank."""
    buf = C.create_string_buffer(128)
    _check(lib().b200r_dist_unique_id(buf))
    return buf.raw


class Pipeline:
    """b200r_pipeline: frames in flight on one rank of `world` (include/b200render.h (1b)). Rank r renders rows r, r+world, ...;
    the rows are assembled into a scan-order frame on every rank by one NCCL all-gather (ASSEMBLE_NCCL) or by peer stores over
    NVLink (ASSEMBLE_PUSH)."""

    def __init__(self, renderer, width, height, depth=2, rank=0, world=1, unique_id=None, assemble=ASSEMBLE_PUSH):
        self._r = renderer
        self._p = C.c_void_p()
        self.depth, self.rank, self.world, self.width, self.height = depth, rank, world, width, height
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(lib().b200r_pipeline_create(renderer._ctx, width, height, depth, rank, world, uid, assemble, C.byref(self._p)))
        self._keep = []
        self.submitted = 0

    def _chk(self, rc):
        if rc != 0:
            raise RendererError((lib().b200r_pipeline_last_error(self._p) or b"").decode() or f"error {rc}")

    def submit(self, frame, host=None):
        """Enqueue a frame; returns its slot. `host`: page-locked numpy array (height x width uint32) or an integer address."""
        ptr = None
        if host is not None:
            ptr = host if isinstance(host, int) else host.ctypes.data
            self._keep = (self._keep + [host])[-(self.depth + 1):]
        self._chk(lib().b200r_pipeline_submit(self._p, C.byref(frame), ptr))
        self.submitted += 1
        return (self.submitted - 1) % self.depth

    def drain(self):
        self._chk(lib().b200r_pipeline_drain(self._p))

    def slot_frame(self, slot):
        """Device address of the assembled frame of `slot` (valid after drain())."""
        ptr = C.c_void_p()
        self._chk(lib().b200r_pipeline_slot_frame(self._p, slot, C.byref(ptr)))
        return ptr.value

    def fence(self, stream, pipeline_waits):
        self._chk(lib().b200r_pipeline_fence(self._p, C.c_void_p(stream), 1 if pipeline_waits else 0))

    def set_l2_flush(self, nbytes, prefetch_scene=True):
        """Measurement aid: evict the L2 (a write of `nbytes`) before every frame, then prefetch the scene back."""
        self._chk(lib().b200r_pipeline_set_l2_flush(self._p, int(nbytes)))
        for k in range(3):
            ptr, n = C.c_void_p(), C.c_uint64()
            if nbytes and prefetch_scene:
                _check(lib().b200r_scene_buffer(self._r._ctx, k, C.byref(ptr), C.byref(n)), self._r._ctx)
            self._chk(lib().b200r_pipeline_set_prefetch(self._p, k, ptr, n.value))

    def launches(self, reset=False):
        n = C.c_uint32()
        self._chk(lib().b200r_pipeline_launches(self._p, C.byref(n), 1 if reset else 0))
        return n.value

    def set_timing(self, enabled):
        self._chk(lib().b200r_pipeline_set_timing(self._p, 1 if enabled else 0))

    def kernel_ms(self):
        """(sum of this rank's render-kernel durations in ms, frames) since the last call; waits for the frames."""
        s, n = C.c_double(), C.c_uint32()
        self._chk(lib().b200r_pipeline_kernel_ms(self._p, C.byref(s), C.byref(n)))
        return s.value, n.value

    def close(self):
        if self._p:
            lib().b200r_pipeline_destroy(self._p)
            self._p = C.c_void_p()


def default_light_pos(index=0):
    p = (C.c_float * 3)()
    lib().b200r_default_light_pos(index, p)
    return tuple(p)
