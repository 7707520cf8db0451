"""ctypes view of the C ABI in include/imba_b200.h (the same binding a C++ host links against).

This module is plumbing for the Python-side tests and bench.py; the product is libimba_b200.so.
Loading fails loudly when the library has not been built: there is no Python or CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libimba_b200.so")


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Mask(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset", C.c_int32)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("diffuse_tex", C.c_int32), ("bump_tex", C.c_int32), ("is_emitter", C.c_int32),
                ("kd", C.c_float * 3), ("eta", C.c_float), ("ks", C.c_float * 3), ("kappa", C.c_float),
                ("tf", C.c_float * 3), ("exponent", C.c_float), ("emit", C.c_float * 3), ("emit_area", C.c_float)]


class Light(C.Structure):
    _fields_ = [("kind", C.c_int32), ("angle", C.c_float), ("pad0", C.c_float * 2), ("intensity", C.c_float * 4),
                ("p0", C.c_float * 4), ("p1", C.c_float * 4), ("p2", C.c_float * 4)]


class Instance(C.Structure):
    _fields_ = [("mesh_id", C.c_int32), ("mat", C.c_float * 12), ("inv_mat", C.c_float * 12)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset", C.c_int64)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("mesh_count", C.c_int32), ("mesh_vertex_base", C.c_void_p), ("mesh_tri_base", C.c_void_p),
        ("vertices", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p), ("geom_normals", C.c_void_p),
        ("indices", C.c_void_p),
        ("instance_count", C.c_int32), ("instances", C.c_void_p),
        ("material_count", C.c_int32), ("materials", C.c_void_p),
        ("light_count", C.c_int32), ("lights", C.c_void_p),
        ("texture_count", C.c_int32), ("textures", C.c_void_p), ("texels", C.c_void_p), ("texel_count", C.c_int64),
        ("env_tex", C.c_int32), ("env_intensity", C.c_float),
        ("bsphere_center", C.c_float * 3), ("bsphere_radius", C.c_float),
        ("accel_layout", C.c_int32), ("root", C.c_int32), ("nodes", C.c_void_p), ("node_count", C.c_int32),
        ("tris", C.c_void_p), ("tri_vec4_count", C.c_int64),
        ("inst_nodes", C.c_void_p), ("inst_node_count", C.c_int32),
        ("mask_count", C.c_int32), ("masks", C.c_void_p), ("mask_buffer", C.c_void_p), ("mask_buffer_size", C.c_int64),
        ("mask_indices", C.c_void_p), ("mask_texcoords", C.c_void_p), ("mask_texcoord_count", C.c_int64),
    ]


class RenderSettings(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("fov", C.c_float),
                ("cam_pos", C.c_float * 3), ("cam_dir", C.c_float * 3), ("cam_up", C.c_float * 3),
                ("spp", C.c_int32), ("max_path_len", C.c_int32), ("light_path_count", C.c_int32),
                ("num_connections", C.c_int32), ("num_knn", C.c_int32), ("radius_factor", C.c_float),
                ("seed", C.c_uint64), ("rank", C.c_int32), ("rank_count", C.c_int32), ("iterations_per_pass", C.c_int32), ("reserved", C.c_int32)]


ALGO = {"pt": 0, "bpt": 1, "vcm": 2, "ppm": 3, "vcm_pt": 4, "lt": 5}

# numpy dtypes of the ABI records
RAY = np.dtype([("org", "<f4", 3), ("tmin", "<f4"), ("dir", "<f4", 3), ("tmax", "<f4")])
HIT = np.dtype([("tri_id", "<i4"), ("inst_id", "<i4"), ("tmax", "<f4"), ("u", "<f4")])
NODE4 = np.dtype([("min_x", "<f4", 4), ("min_y", "<f4", 4), ("min_z", "<f4", 4), ("max_x", "<f4", 4), ("max_y", "<f4", 4),
                  ("max_z", "<f4", 4), ("children", "<i4", 4)])
NODE2 = np.dtype([("left_bb", "<f4", 6), ("right_bb", "<f4", 6), ("left", "<i4"), ("right", "<i4")])
INSTANCE_NODE = np.dtype([("transf", "<f4", (3, 4)), ("id", "<i4"), ("next", "<i4"), ("pad", "<i4", 2)])
MATERIAL = np.dtype([("kind", "<i4"), ("diffuse_tex", "<i4"), ("bump_tex", "<i4"), ("is_emitter", "<i4"), ("kd", "<f4", 3),
                     ("eta", "<f4"), ("ks", "<f4", 3), ("kappa", "<f4"), ("tf", "<f4", 3), ("exponent", "<f4"),
                     ("emit", "<f4", 3), ("emit_area", "<f4")])
LIGHT = np.dtype([("kind", "<i4"), ("angle", "<f4"), ("pad0", "<f4", 2), ("intensity", "<f4", 4), ("p0", "<f4", 4),
                  ("p1", "<f4", 4), ("p2", "<f4", 4)])
INSTANCE = np.dtype([("mesh_id", "<i4"), ("mat", "<f4", 12), ("inv_mat", "<f4", 12)])
TEXTURE = np.dtype([("width", "<i4"), ("height", "<i4"), ("offset", "<i8")])
MASK = np.dtype([("width", "<i4"), ("height", "<i4"), ("offset", "<i4")])
NODE8 = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"), ("prim_base", "<u4"),
                  ("meta", "u1", 8), ("q", "u1", (6, 8))])
PT_STATE = np.dtype([("pixel_id", "<i4"), ("sample_id", "<i4"), ("rng", "<u8"), ("throughput", "<f4", 3), ("bounces_spec", "<u4"),
                     ("last_pdf", "<f4"), ("pad", "<u4")])
VCM_STATE = np.dtype([("pixel_id", "<i4"), ("light_id", "<i4"), ("rng", "<u8"), ("throughput", "<f4", 3), ("len_finite", "<u4"),
                      ("dVC", "<f4"), ("dVCM", "<f4"), ("dVM", "<f4"), ("pad", "<u4")])
LIGHT_VERTEX = np.dtype([("pos", "<f4", 3), ("path_length", "<i4"), ("out_dir", "<f4", 3), ("mat", "<i4"), ("normal", "<f4", 3), ("dVC", "<f4"),
                         ("geom_normal", "<f4", 3), ("dVCM", "<f4"), ("u_tangent", "<f4", 3), ("dVM", "<f4"), ("v_tangent", "<f4", 3), ("uv_x", "<f4"),
                         ("throughput", "<f4", 3), ("uv_y", "<f4")])
assert VCM_STATE.itemsize == 48 and LIGHT_VERTEX.itemsize == 112
SHADOW_STATE = np.dtype([("pixel_id", "<i4"), ("throughput", "<f4", 3)])
assert PT_STATE.itemsize == 40 and SHADOW_STATE.itemsize == 16
TRI48 = np.dtype([("v0", "<f4", 3), ("id", "<i4"), ("e1", "<f4", 3), ("nx", "<f4"), ("e2", "<f4", 3), ("ny", "<f4")])
INST64 = np.dtype([("m", "<f4", 12), ("inst_id", "<i4"), ("root", "<i4"), ("pad", "<i4", 2)])
assert RAY.itemsize == 32 and HIT.itemsize == 16 and NODE4.itemsize == 112 and NODE2.itemsize == 56
assert INSTANCE_NODE.itemsize == 64 and MATERIAL.itemsize == C.sizeof(Material) == 80 and LIGHT.itemsize == C.sizeof(Light) == 80
assert INSTANCE.itemsize == C.sizeof(Instance) == 100 and TEXTURE.itemsize == C.sizeof(Texture) == 16
assert NODE8.itemsize == 80 and TRI48.itemsize == 48 and INST64.itemsize == 64

_lib = None


def lib():
    """Returns the loaded libimba_b200.so with argument types declared.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C imbatracer_b200/csrc).  There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.imba_last_error.restype = C.c_char_p
    L.imba_launch_count.restype = u64
    L.imba_scene_create.restype = vp; L.imba_scene_create.argtypes = [C.POINTER(SceneDesc)]
    L.imba_scene_destroy.restype = None; L.imba_scene_destroy.argtypes = [vp]
    L.imba_scene_stats.argtypes = [vp, C.POINTER(i64 * 5)]
    L.imba_scene_build_info.argtypes = [vp, C.POINTER(C.c_float * 2)]
    L.imba_trace.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.imba_trace_host.argtypes = [vp, vp, vp, i32, i32]
    L.imba_host_register.argtypes = [vp, C.c_size_t]
    L.imba_host_unregister.argtypes = [vp]
    L.imba_bvh8_build.restype = vp; L.imba_bvh8_build.argtypes = [i32, i32, vp, vp, vp]
    L.imba_bvh8_free.restype = None; L.imba_bvh8_free.argtypes = [vp]
    L.imba_bvh8_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64), C.POINTER(vp),
                                   C.POINTER(i64), C.POINTER(C.c_uint32), C.POINTER(i32)]
    f32, u32 = C.c_float, C.c_uint32
    L.imba_bvh_cache_open.restype = vp; L.imba_bvh_cache_open.argtypes = [C.c_char_p]
    L.imba_bvh_cache_close.restype = None; L.imba_bvh_cache_close.argtypes = [vp]
    L.imba_bvh_cache_block.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u32), C.POINTER(vp), C.POINTER(u32)]
    L.imba_bvh_cache_append.argtypes = [C.c_char_p, i32, vp, u32, vp, u32]
    L.imba_accel_build.restype = vp; L.imba_accel_build.argtypes = [C.POINTER(SceneDesc)]
    L.imba_accel_build_gpu.restype = vp; L.imba_accel_build_gpu.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_float)]
    L.imba_accel_free.restype = None; L.imba_accel_free.argtypes = [vp]
    L.imba_accel_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64), C.POINTER(i32)]
    L.imba_integrator_create.restype = vp; L.imba_integrator_create.argtypes = [vp, C.POINTER(RenderSettings)]
    L.imba_integrator_destroy.restype = None; L.imba_integrator_destroy.argtypes = [vp]
    L.imba_integrator_preprocess.argtypes = [vp]
    L.imba_integrator_pixel_size.restype = f32; L.imba_integrator_pixel_size.argtypes = [vp]
    L.imba_integrator_vertex_cache_size.argtypes = [vp]
    L.imba_integrator_reset.argtypes = [vp]
    L.imba_integrator_set_camera.argtypes = [vp, vp, vp, vp]
    L.imba_integrator_allreduce.argtypes = [vp, vp, vp]
    L.imba_queue_compact_hits_host.argtypes = [vp, vp, vp, i32, i32, vp]
    L.imba_queue_compact_rays_host.argtypes = [vp, vp, i32, i32, vp]
    L.imba_queue_sort_keys_host.argtypes = [vp, i32, i32, vp]
    L.imba_integrator_render.argtypes = [vp, i32]
    L.imba_integrator_framebuffer.restype = vp; L.imba_integrator_framebuffer.argtypes = [vp]
    L.imba_integrator_set_framebuffer.argtypes = [vp, vp]
    L.imba_integrator_stream.restype = vp; L.imba_integrator_stream.argtypes = [vp]
    L.imba_integrator_sync.argtypes = [vp]
    L.imba_integrator_download.argtypes = [vp, vp]
    L.imba_integrator_clear.argtypes = [vp]
    L.imba_integrator_ray_counts.argtypes = [vp, vp]
    L.imba_integrator_camera_rays.argtypes = [vp, i32, vp, vp]
    L.imba_seed_states.argtypes = [u32, i32, i32, i32, vp]
    L.imba_pt_shade_batch.argtypes = [vp, i32] + [vp] * 9
    L.imba_vcm_stage_batch.argtypes = [vp, i32, i32, f32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.imba_vcm_raygen_batch.argtypes = [vp, i32, i32, f32, u32, i32, vp, vp]
    L.imba_queue_sort_by_material.argtypes = [vp, vp, i32, vp, vp, vp]
    L.imba_queue_compact_rays.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp]
    L.imba_queue_compact_hits.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp]
    _lib = L
    return L


def last_error():
    return lib().imba_last_error().decode()


def check(rc, what="call"):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")
