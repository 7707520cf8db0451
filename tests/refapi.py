"""TEST INFRASTRUCTURE: ctypes bindings of the oracle (oracle/liboracle.so) and of the reference host code compiled
in place (oracle/_ref/libimba_ref.so).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module; nothing in imbatracer_b200/ does."""
import ctypes as C
import os
import shutil

import numpy as np

from imbatracer_b200 import abi
from imbatracer_b200.scene import FlatScene, from_desc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libimba_ref.so")
REF_B200_SO = os.path.join(ROOT, "oracle", "_ref", "libimba_ref_b200.so")   # the same reference sources linked against the product library
REF_SCENES = os.path.join(ROOT, "oracle", "_ref", "scenes")  # scratch COPY of the reference's test/scenes (git-ignored)

_oracle = None
_ref = None

ENTRY = {  # (layout, any_hit) -> symbol
    (0, 0): "intersect_cpu_masked_instanced", (0, 1): "occluded_cpu_masked_instanced",
    (1, 0): "intersect_gpu_masked_instanced", (1, 1): "occluded_gpu_masked_instanced",
}


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(ORACLE_SO)
        vp, i32 = C.c_void_p, C.c_int32
        for name in ENTRY.values():
            getattr(L, name).argtypes = [i32] + [vp] * 9 + [i32]
            getattr(L, name).restype = None
            getattr(L, name + "_counted").argtypes = [i32] + [vp] * 9 + [i32, vp]
            getattr(L, name + "_counted").restype = None
        L.oracle_set_threads.argtypes = [i32]
        L.oracle_get_threads.restype = i32
        L.bvh8_walk.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, i32, i32, vp]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_b200():
    return os.path.exists(REF_B200_SO)


_ref_b200 = None


def ref_b200():
    """oracle/_ref/libimba_ref_b200.so (oracle/Makefile: ref_b200): the unmodified reference whose traversal symbols resolve to
    imbatracer_b200/libimba_b200.so and whose AnyDSL arrays live in CUDA memory.  Needs a GPU."""
    global _ref_b200
    if _ref_b200 is None:
        L = C.CDLL(REF_B200_SO)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.ref_scene_load.restype = vp; L.ref_scene_load.argtypes = [C.c_char_p, i32]
        L.ref_scene_free.restype = None; L.ref_scene_free.argtypes = [vp]
        L.ref_render.argtypes = [vp, i32, i32, i32, f32] + [i32] * 8 + [f32, vp, vp, vp, vp]
        L.ref_set_seed.restype = None; L.ref_set_seed.argtypes = [C.c_uint64]
        L.ref_render_plugin.argtypes = [vp, i32, i32, i32, i32, f32] + [i32] * 8 + [f32, C.c_uint32, vp, vp, vp]
        assert L.ref_is_linked_to_product() == 1
        _ref_b200 = L
    return _ref_b200


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.ref_scene_load.restype = vp; L.ref_scene_load.argtypes = [C.c_char_p, i32]
        L.ref_scene_free.restype = None; L.ref_scene_free.argtypes = [vp]
        L.ref_scene_camera.argtypes = [vp, vp, vp, vp]
        L.ref_scene_export.argtypes = [vp, C.POINTER(abi.SceneDesc)]
        L.ref_render.argtypes = [vp, i32, i32, i32, f32] + [i32] * 8 + [f32, vp, vp, vp, vp]
        L.ref_camera_rays.argtypes = [vp, i32, i32, f32, vp, vp]
        L.ref_pt_shade_batch.argtypes = [vp, i32, i32, f32, i32, i32] + [vp] * 10
        L.ref_calc_intersections.argtypes = [vp, i32, vp, vp, vp, vp]
        L.ref_rng_floats.restype = None; L.ref_rng_floats.argtypes = [C.c_uint64, i32, i32, vp, vp]
        L.ref_seed_states.restype = None; L.ref_seed_states.argtypes = [C.c_uint64, i32, i32, i32, vp]
        L.ref_struct_sizes.argtypes = [vp]
        L.ref_vcm_light_emit.argtypes = [vp, i32, vp, vp, f32, vp, vp]
        L.ref_vcm_camera_start.argtypes = [vp, i32, i32, f32, i32, i32, vp, vp, vp, vp]
        L.ref_scene_add_test_textures.argtypes = [vp, C.c_uint64]
        L.ref_scene_add_test_envmap.argtypes = [vp, C.c_uint64, f32]
        L.ref_scene_add_test_phong.argtypes = [vp, C.c_uint64]
        L.ref_set_seed.restype = None; L.ref_set_seed.argtypes = [C.c_uint64]
        L.ref_vcm_cache_size.argtypes = [vp, i32]
        L.ref_vcm_stage.argtypes = [vp, i32, i32, i32, f32, i32, i32, i32, i32, f32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        _ref = L
    return _ref


def trace_oracle(scene: FlatScene, rays: np.ndarray, any_hit=False, counters=False, threads=0):
    """Runs the CPU restatement on the reference-layout structure held by `scene`. Returns hits (and counters)."""
    L = oracle()
    L.oracle_set_threads(threads)
    rays = np.ascontiguousarray(rays); assert rays.dtype == abi.RAY
    hits = np.zeros(len(rays), abi.HIT)
    a = scene.arrays
    masks = a["masks"].ctypes.data if len(a["masks"]) else None
    args = [scene.root, a["nodes"].ctypes.data, a["inst_nodes"].ctypes.data, a["tris"].ctypes.data, rays.ctypes.data,
            hits.ctypes.data, a["mask_indices"].ctypes.data if masks else None, a["mask_texcoords"].ctypes.data if masks else None,
            masks, a["mask_buffer"].ctypes.data if masks else None, len(rays)]
    name = ENTRY[(scene.accel_layout, int(bool(any_hit)))]
    if counters:
        cnt = np.zeros((len(rays), 4), np.uint32)
        getattr(L, name + "_counted")(*args, cnt.ctypes.data)
        return hits, cnt
    getattr(L, name)(*args)
    return hits


class Bvh8Host:
    """Host arrays of the product's converted structure (imba_bvh8_build) -- for the CPU walker."""

    def __init__(self, scene: FlatScene):
        L = abi.lib()
        a = scene.arrays
        self.h = L.imba_bvh8_build(scene.accel_layout, scene.root, a["nodes"].ctypes.data, a["tris"].ctypes.data, a["inst_nodes"].ctypes.data)
        if not self.h:
            raise RuntimeError(abi.last_error())
        p = [C.c_void_p() for _ in range(3)]; n = [C.c_int64() for _ in range(3)]; root = C.c_uint32(); depth = C.c_int32()
        abi.check(L.imba_bvh8_arrays(self.h, p[0], n[0], p[1], n[1], p[2], n[2], root, depth))

        def grab(ptr, count, dt):
            buf = (C.c_uint8 * (count * dt.itemsize)).from_address(ptr.value)
            return np.frombuffer(buf, dtype=dt, count=count).copy()

        self.nodes = grab(p[0], n[0].value, abi.NODE8); self.tris = grab(p[1], n[1].value, abi.TRI48)
        self.insts = grab(p[2], n[2].value, abi.INST64); self.top_root = root.value; self.max_depth = depth.value
        L.imba_bvh8_free(self.h); self.h = None

    def walk(self, rays, any_hit=False, stats=False):
        rays = np.ascontiguousarray(rays); hits = np.zeros(len(rays), abi.HIT)
        st = np.zeros((len(rays), 4), np.uint32) if stats else None
        oracle().bvh8_walk(self.nodes.ctypes.data, self.tris.ctypes.data, self.insts.ctypes.data, self.top_root,
                           rays.ctypes.data, hits.ctypes.data, len(rays), int(bool(any_hit)), st.ctypes.data if stats else None)
        return (hits, st) if stats else hits


def ensure_ref_scenes(reference_root="/root/reference"):
    """Copies the reference's test scenes to a scratch dir (its loader writes .bvh caches next to them, scene.cpp:67)."""
    if not os.path.isdir(REF_SCENES):
        src = os.path.join(reference_root, "test", "scenes", "cornell")
        if not os.path.isdir(src):
            return None
        os.makedirs(REF_SCENES, exist_ok=True)
        shutil.copytree(src, os.path.join(REF_SCENES, "cornell"))
    ensure_proxy_scene()
    return REF_SCENES


PROXY_DETAIL = 0.02   # ~31 k triangles: what the parity fixtures and the reference-side harness use


def ensure_proxy_scene():
    """Writes the procedural Still Life stand-in (imbatracer_b200/procedural.py) as OBJ/MTL/TGA/.scene into the scratch
    scene directory, for the reference's loader."""
    d = os.path.join(REF_SCENES, "stilllife_proxy")
    if not os.path.exists(os.path.join(d, "stilllife_proxy.scene")):
        from imbatracer_b200 import procedural
        from tests import obj_export
        obj_export.write_scene(procedural.still_life_proxy(PROXY_DETAIL, accel=None), d)
    return d


def scene_path(name):
    """.scene file of a test scene inside the scratch copy."""
    if name == "stilllife_proxy":
        return os.path.join(ensure_proxy_scene(), "stilllife_proxy.scene")
    return os.path.join(REF_SCENES, "cornell", name + ".scene")


class RefScene:
    """A scene loaded by the reference's own loader + SBVH builder + adapters."""

    def __init__(self, scene_file, gpu_layout=False, lib=None):
        self.L = lib or ref()
        self.h = self.L.ref_scene_load(scene_file.encode(), int(gpu_layout))
        if not self.h:
            raise RuntimeError(f"reference loader failed on {scene_file}")
        self.gpu_layout = gpu_layout

    def close(self):
        if self.h:
            self.L.ref_scene_free(self.h); self.h = None

    def add_test_textures(self, seed=99):
        """Procedural colour / bump textures + texcoords on the reference side (oracle/ref_driver.cpp); export with flat() afterwards."""
        return self.L.ref_scene_add_test_textures(self.h, seed)

    def add_test_phong(self, seed=17):
        """Replaces the diffuse / glossy materials by materials that build the reference's Phong BxDF (brdfs.h:48-92); export with flat()."""
        return self.L.ref_scene_add_test_phong(self.h, seed)

    def add_test_envmap(self, seed=5, intensity=0.8):
        """Procedural lat-long environment map + EnvLight on the reference side; export with flat() afterwards."""
        return self.L.ref_scene_add_test_envmap(self.h, seed, intensity)

    def flat(self, name="") -> FlatScene:
        d = abi.SceneDesc()
        rc = self.L.ref_scene_export(self.h, C.byref(d))
        if rc != 0:
            raise RuntimeError(f"ref_scene_export failed: {rc}")
        s = from_desc(d, 1 if self.gpu_layout else 0)
        cam = np.zeros(9, np.float32)
        self.L.ref_scene_camera(self.h, cam[0:3].ctypes.data, cam[3:6].ctypes.data, cam[6:9].ctypes.data)
        s.cam = cam; s.name = name
        return s

    def camera_rays(self, w, h, fov=60.0, jitter=None):
        rays = np.zeros(w * h, abi.RAY)
        j = None if jitter is None else np.ascontiguousarray(jitter, np.float32)
        self.L.ref_camera_rays(self.h, w, h, fov, j.ctypes.data if j is not None else None, rays.ctypes.data)
        return rays

    def render(self, algo, w, h, fov=60.0, spp=1, iterations=1, threads=0, tile=256, max_path_len=10, light_paths=0,
               connections=1, knn=10, radius_factor=2.0):
        img = np.zeros((h, w, 3), np.float32); stats = np.zeros(2, np.uint64); secs = C.c_double(); px = C.c_float()
        threads = threads or os.cpu_count()
        rc = self.L.ref_render(self.h, algo, w, h, fov, spp, iterations, threads, tile, max_path_len, light_paths, connections, knn,
                               radius_factor, img.ctypes.data, stats.ctypes.data, C.byref(secs), C.byref(px))
        assert rc == 0
        return img, {"primary_rays": int(stats[0]), "shadow_rays": int(stats[1]), "seconds": secs.value, "pixel_size": px.value}

    def render_plugin(self, mode, algo, w, h, fov=60.0, spp=1, iterations=1, threads=0, tile=256, max_path_len=10, light_paths=0, connections=1,
                      knn=10, radius_factor=2.0, seed=1):
        """libimba_ref_b200.so only: mode 0 = the reference's integrators on imba::B200WavefrontScheduler, mode 1 = imba::B200Integrator
        (include/imba_b200_ref.hpp), both driven through the reference's Integrator interface."""
        img = np.zeros((h, w, 3), np.float32); stats = np.zeros(2, np.uint64); px = np.zeros(2, np.float32)
        rc = self.L.ref_render_plugin(self.h, mode, algo, w, h, fov, spp, iterations, threads or os.cpu_count(), tile, max_path_len, light_paths, connections, knn,
                                      radius_factor, seed, img.ctypes.data, stats.ctypes.data, px.ctypes.data)
        assert rc == 0, rc
        return img, {"primary_rays": int(stats[0]), "shadow_rays": int(stats[1]), "pixel_size": float(px[0]), "device_pixel_size": float(px[1])}

    def vcm_cache_size(self, light_path_count):
        """LightVertices::compute_cache_size of the reference (light_vertices.cpp:15-129)."""
        return int(self.L.ref_vcm_cache_size(self.h, light_path_count))

    def light_emit(self, light_ids, rng_states, mis_eta_vc):
        """Light::sample_emit + initial VCM weights of the reference for (light, RNG state) pairs (ref_driver.cpp: ref_vcm_light_emit)."""
        n = len(light_ids)
        light_ids = np.ascontiguousarray(light_ids, np.int32); rng_states = np.ascontiguousarray(rng_states, np.uint64)
        rays = np.zeros(n, abi.RAY); states = np.zeros(n, abi.VCM_STATE)
        rc = self.L.ref_vcm_light_emit(self.h, n, light_ids.ctypes.data, rng_states.ctypes.data, mis_eta_vc, rays.ctypes.data, states.ctypes.data)
        assert rc == 0
        return rays, states

    def camera_start(self, w, h, light_path_count, pixel_ids, rng_states, fov=60.0):
        """Camera-path start of the reference's VCM integrator for (pixel, RNG state) pairs (ref_driver.cpp: ref_vcm_camera_start)."""
        n = len(pixel_ids)
        pixel_ids = np.ascontiguousarray(pixel_ids, np.int32); rng_states = np.ascontiguousarray(rng_states, np.uint64)
        rays = np.zeros(n, abi.RAY); states = np.zeros(n, abi.VCM_STATE)
        rc = self.L.ref_vcm_camera_start(self.h, w, h, fov, light_path_count, n, pixel_ids.ctypes.data, rng_states.ctypes.data, rays.ctypes.data, states.ctypes.data)
        assert rc == 0
        return rays, states

    def pt_shade_batch(self, w, h, rays, hits, states, fov=60.0, max_path_len=10):
        """PathTracer::process_primary_rays (pt.cpp:82-160) of the reference on an injected batch."""
        n = len(rays)
        rays_out = np.zeros(n, abi.RAY); states_out = np.zeros(n, abi.PT_STATE)
        sh_rays = np.zeros(n, abi.RAY); sh_states = np.zeros(n, abi.SHADOW_STATE)
        fb = np.zeros((h, w, 3), np.float32)
        nc, ns = C.c_int(), C.c_int()
        rc = self.L.ref_pt_shade_batch(self.h, w, h, fov, max_path_len, n, rays.ctypes.data, hits.ctypes.data, states.ctypes.data,
                                       rays_out.ctypes.data, states_out.ctypes.data, C.addressof(nc), sh_rays.ctypes.data,
                                       sh_states.ctypes.data, C.addressof(ns), fb.ctypes.data)
        assert rc == 0
        return rays_out[:nc.value], states_out[:nc.value], sh_rays[:ns.value], sh_states[:ns.value], fb

    def vcm_stage(self, algo, w, h, stage, iteration, base_radius, rays, hits, states, cache_in=None, light_path_count=0, max_path_len=10,
                  num_connections=1, num_knn=10, fov=60.0):
        """VCMIntegrator<algo>::process_light_rays (stage 0) / process_camera_rays (stage 1) of the reference on a batch."""
        n = len(rays); slots = num_connections + 1
        rays_out = np.zeros(n, abi.RAY); states_out = np.zeros(n, abi.VCM_STATE)
        sh_rays = np.zeros(n * slots, abi.RAY); sh_states = np.zeros(n * slots, abi.SHADOW_STATE)
        fb = np.zeros((h, w, 3), np.float32); cache_out = np.zeros(2 * n + 16, abi.LIGHT_VERTEX)
        nc, ns, nv = C.c_int(), C.c_int(), C.c_int()
        cin = np.ascontiguousarray(cache_in) if cache_in is not None else np.zeros(0, abi.LIGHT_VERTEX)
        rc = self.L.ref_vcm_stage(self.h, algo, w, h, fov, max_path_len, light_path_count or w * h // 2, num_connections, num_knn, base_radius, iteration, stage, n,
                                  rays.ctypes.data, hits.ctypes.data, states.ctypes.data, cin.ctypes.data if len(cin) else None, len(cin),
                                  rays_out.ctypes.data, states_out.ctypes.data, C.addressof(nc), sh_rays.ctypes.data, sh_states.ctypes.data, C.addressof(ns),
                                  fb.ctypes.data, cache_out.ctypes.data, C.addressof(nv))
        assert rc == 0
        return rays_out[:nc.value], states_out[:nc.value], sh_rays[:ns.value], sh_states[:ns.value], fb, cache_out[:nv.value]
