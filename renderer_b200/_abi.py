"""ctypes mirror of include/b200render.h (plain structs and prototypes; no torch types)."""
import ctypes as C

MAX_LIGHTS = 2
MAX_FRAMES_IN_FLIGHT = 16      # B200R_MAX_FRAMES_IN_FLIGHT
SHADOWMAP_SIZE = 1024

F_SHADOWS, F_REFLECTIONS, F_PHONG_NORMAL, F_AO, F_MLAA = 0x01, 0x02, 0x04, 0x08, 0x10
F_DEFAULT = F_SHADOWS | F_REFLECTIONS | F_PHONG_NORMAL

(MODE_POINTS, MODE_POINTS_TRI, MODE_LINES, MODE_AMBIENT, MODE_GOURAUD, MODE_PHONG,
 MODE_PHONG_SHADOWMAPS, MODE_PHONG_SOFTSHADOWMAPS, MODE_RAYTRACE, MODE_RAYTRACE_AA) = range(1, 11)


class Vertex(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("nrm", C.c_float * 3), ("ao", C.c_uint32)]


class Tri(C.Structure):
    _fields_ = [("a", C.c_uint32), ("b", C.c_uint32), ("c", C.c_uint32),
                ("center", C.c_float * 3), ("normal", C.c_float * 3), ("colorf", C.c_float * 3),
                ("color", C.c_uint32), ("two_sided", C.c_uint32),
                ("d", C.c_float), ("d1", C.c_float), ("d2", C.c_float), ("d3", C.c_float),
                ("e1", C.c_float * 3), ("e2", C.c_float * 3), ("e3", C.c_float * 3)]


class BvhNode(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("hi", C.c_float * 3), ("a", C.c_uint32), ("b", C.c_uint32)]


class Light(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("in_camera", C.c_float * 3), ("cam2light", C.c_float * 9)]


class Frame(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("eye", C.c_float * 3), ("mv", C.c_float * 9),
                ("n_lights", C.c_uint32), ("lights", Light * MAX_LIGHTS),
                ("flags", C.c_uint32), ("ao_samples", C.c_uint32), ("max_depth", C.c_uint32),
                ("frame_index", C.c_uint32), ("row_first", C.c_uint32), ("row_step", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rays_primary", "rays_shadow", "rays_reflection", "rays_ao",
        "node_tests", "leaf_visits", "tri_tests", "tris_setup", "spans", "z_tests", "z_passes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Orbit(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("angle1", C.c_float), ("angle2", C.c_float), ("d_angle", C.c_float)]


assert C.sizeof(Vertex) == 28 and C.sizeof(BvhNode) == 32 and C.sizeof(Tri) == 108

# every symbol include/b200render.h declares: name -> (restype, argtypes)
P = C.POINTER
SYMBOLS = {
    "b200r_init": (C.c_int, [C.c_int, P(C.c_void_p)]),
    "b200r_destroy": (None, [C.c_void_p]),
    "b200r_last_error": (C.c_char_p, [C.c_void_p]),
    "b200r_upload_scene": (C.c_int, [C.c_void_p, P(Vertex), C.c_uint32, P(Tri), C.c_uint32,
                                     P(BvhNode), C.c_uint32, P(C.c_int32), C.c_uint32]),
    "b200r_upload_shadowmap": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200r_render_shadowmap": (C.c_int, [C.c_void_p, C.c_int, P(C.c_float), P(C.c_float)]),
    "b200r_download_shadowmap": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200r_render": (C.c_int, [C.c_void_p, P(Frame), C.c_void_p]),
    "b200r_render_async": (C.c_int, [C.c_void_p, P(Frame), C.c_void_p]),
    "b200r_wait": (C.c_int, [C.c_void_p]),
    "b200r_render_device": (C.c_int, [C.c_void_p, P(Frame), C.c_void_p, C.c_void_p]),
    "b200r_render_device_slot": (C.c_int, [C.c_void_p, P(Frame), C.c_void_p, C.c_void_p, C.c_uint32]),
    "b200r_set_pipeline_depth": (C.c_int, [C.c_void_p, C.c_uint32]),
    "b200r_mlaa_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "b200r_deinterleave_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_void_p]),
    "b200r_selftest_division": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, P(C.c_uint64), P(C.c_float)]),
    "b200r_build_bvh": (C.c_int, [C.c_void_p, P(Vertex), C.c_uint32, P(Tri), C.c_uint32, P(BvhNode), C.c_uint32,
                                  P(C.c_int32), P(C.c_uint32), P(C.c_int32)]),
    "b200r_selftest_bvh_steps_host": (C.c_int, [P(Vertex), C.c_uint32, P(Tri), C.c_uint32, P(BvhNode), C.c_uint32,
                                                P(C.c_int32), P(C.c_uint32), P(C.c_int32)]),
    "b200r_selftest_span_walk_host": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_uint64)]),
    "b200r_selftest_mlaa_steps_host": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]),
    "b200r_set_tile_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "b200r_get_tile_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, P(C.c_uint32)]),
    "b200r_set_switch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "b200r_dist_unique_id": (C.c_int, [C.c_void_p]),
    "b200r_pipeline_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                        P(C.c_void_p)]),
    "b200r_pipeline_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200r_pipeline_drain": (C.c_int, [C.c_void_p]),
    "b200r_pipeline_slot_frame": (C.c_int, [C.c_void_p, C.c_uint32, P(C.c_void_p)]),
    "b200r_pipeline_fence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200r_pipeline_set_l2_flush": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200r_pipeline_set_prefetch": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "b200r_pipeline_launches": (C.c_int, [C.c_void_p, P(C.c_uint32), C.c_int]),
    "b200r_pipeline_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "b200r_pipeline_kernel_ms": (C.c_int, [C.c_void_p, P(C.c_double), P(C.c_uint32)]),
    "b200r_pipeline_last_error": (C.c_char_p, [C.c_void_p]),
    "b200r_pipeline_destroy": (None, [C.c_void_p]),
    "b200r_host_alloc": (C.c_int, [C.c_uint64, P(C.c_void_p)]),
    "b200r_host_free": (None, [C.c_void_p]),
    "b200r_scene_buffer": (C.c_int, [C.c_void_p, C.c_uint32, P(C.c_void_p), P(C.c_uint64)]),
    "b200r_set_counters": (C.c_int, [C.c_void_p, C.c_int]),
    "b200r_get_counters": (C.c_int, [C.c_void_p, P(Counters)]),
    "b200r_last_kernel_ms": (C.c_int, [C.c_void_p, P(C.c_float), P(C.c_float)]),
    "b200r_last_launches": (C.c_int, [C.c_void_p, P(C.c_uint32)]),
    "b200r_scene_load": (C.c_int, [C.c_char_p, P(C.c_void_p)]),
    "b200r_scene_free": (None, [C.c_void_p]),
    "b200r_scene_build_bvh": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "b200r_scene_build_bvh_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]),
    "b200r_scene_vertices": (P(Vertex), [C.c_void_p, P(C.c_uint32)]),
    "b200r_scene_tris": (P(Tri), [C.c_void_p, P(C.c_uint32)]),
    "b200r_scene_nodes": (P(BvhNode), [C.c_void_p, P(C.c_uint32)]),
    "b200r_scene_tri_idx": (P(C.c_int32), [C.c_void_p, P(C.c_uint32)]),
    "b200r_scene_bvh_depth": (C.c_int, [C.c_void_p]),
    "b200r_scene_unbounded_triangles": (C.c_uint32, [C.c_void_p, C.c_double]),
    "b200r_upload_scene_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200r_camera_look_at": (None, [P(C.c_float), P(C.c_float), P(C.c_float)]),
    "b200r_orbit_init": (None, [P(Orbit)]),
    "b200r_orbit_step": (None, [P(Orbit), P(C.c_float), P(C.c_float)]),
    "b200r_default_light_pos": (None, [C.c_int, P(C.c_float)]),
    "b200r_light_in_camera_space": (None, [P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_float)]),
    "b200r_light_camera_to_light": (None, [P(C.c_float), P(C.c_float), P(C.c_float)]),
    "b200r_light_world_to_light": (None, [P(C.c_float), P(C.c_float)]),
    "b200r_frame_defaults": (None, [P(Frame), C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_float), P(C.c_float),
                                    C.c_uint32]),
    "b200r_version": (C.c_char_p, []),
}


def bind(lib):
    """Attach prototypes; raises AttributeError naming the first missing export."""
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
