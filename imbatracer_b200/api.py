"""Thin Python handle classes over the C ABI (include/imba_b200.h).  torch is used only for device memory and
streams; every computation happens in libimba_b200.so."""
import ctypes as C

import numpy as np

from . import abi
from .scene import FlatScene


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class Scene:
    """imba_scene: the flattened scene on the current CUDA device with the compressed 8-wide BVH."""

    def __init__(self, flat: FlatScene, device_build=False):
        """device_build: ignore any acceleration structure in `flat` and build it on the device from the meshes and
        instances (IMBA_ACCEL_DEVICE_BUILD: linear BVH + 8-wide collapse, csrc/bvh8_gpu.cu)."""
        self.L = abi.lib()
        self.flat = flat
        d = flat.desc()
        if device_build:
            d.accel_layout = 2; d.nodes = None; d.tris = None; d.inst_nodes = None
            d.node_count = 0; d.tri_vec4_count = 0; d.inst_node_count = 0
        self.h = self.L.imba_scene_create(C.byref(d))
        if not self.h:
            raise RuntimeError("imba_scene_create failed: " + abi.last_error())

    def close(self):
        if self.h:
            self.L.imba_scene_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        out = (C.c_int64 * 5)()
        abi.check(self.L.imba_scene_stats(self.h, C.byref(out)), "imba_scene_stats")
        return {"nodes": out[0], "triangles": out[1], "instances": out[2], "bytes": out[3], "max_depth": out[4]}

    def build_info(self):
        out = (C.c_float * 2)()
        abi.check(self.L.imba_scene_build_info(self.h, C.byref(out)), "imba_scene_build_info")
        return {"device_seconds": out[0], "mesh_levels": int(out[1])}

    def trace(self, rays, hits=None, any_hit=False, counters=None, count=None, stream=None):
        """rays: CUDA uint8/float32 tensor holding imba_ray[count]; returns the hits tensor (int32 [count, 4])."""
        import torch
        n = count if count is not None else rays.numel() * rays.element_size() // 32
        if hits is None:
            hits = torch.empty((n, 4), dtype=torch.int32, device=rays.device)
        abi.check(self.L.imba_trace(self.h, rays.data_ptr(), hits.data_ptr(), n, int(any_hit),
                                    counters.data_ptr() if counters is not None else None, _stream_ptr(stream)), "imba_trace")
        return hits

    def trace_host(self, rays: np.ndarray, hits: np.ndarray = None, any_hit=False):
        """rays: numpy array of abi.RAY (host memory, pageable or pinned); synchronous."""
        assert rays.dtype == abi.RAY and rays.flags["C_CONTIGUOUS"]
        if hits is None:
            hits = np.empty(len(rays), abi.HIT)
        abi.check(self.L.imba_trace_host(self.h, rays.ctypes.data, hits.ctypes.data, len(rays), int(any_hit)), "imba_trace_host")
        return hits


def rays_to_device(rays: np.ndarray, device="cuda"):
    import torch
    return torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).to(device)


def hits_from_device(hits) -> np.ndarray:
    return hits.cpu().numpy().reshape(-1, 4).copy().view(abi.HIT).reshape(-1)


def render_settings(flat, algorithm="pt", width=512, height=512, fov=60.0, spp=1, max_path_len=10, light_path_count=0,
                    num_connections=1, num_knn=10, radius_factor=2.0, seed=1, rank=0, rank_count=1, iterations_per_pass=1):
    """UserSettings defaults of the reference (frontend/cmd_line.h:61-79) + the camera of the .scene file."""
    s = abi.RenderSettings()
    s.algorithm = abi.ALGO[algorithm]; s.width = width; s.height = height; s.fov = fov
    s.cam_pos[:] = [float(x) for x in flat.cam[0:3]]; s.cam_dir[:] = [float(x) for x in flat.cam[3:6]]; s.cam_up[:] = [float(x) for x in flat.cam[6:9]]
    s.spp = spp; s.max_path_len = max_path_len; s.light_path_count = light_path_count or width * height // 2
    s.num_connections = num_connections; s.num_knn = num_knn; s.radius_factor = radius_factor
    s.seed = seed; s.rank = rank; s.rank_count = rank_count; s.iterations_per_pass = iterations_per_pass
    return s


class Integrator:
    """imba_integrator: Integrator::{preprocess, render, reset, pixel_size} on the device (integrator.h:19-39)."""

    def __init__(self, scene: Scene, settings: abi.RenderSettings):
        self.L = abi.lib(); self.scene = scene; self.settings = settings
        self.h = self.L.imba_integrator_create(scene.h, C.byref(settings))
        if not self.h:
            raise RuntimeError("imba_integrator_create failed: " + abi.last_error())

    def close(self):
        if self.h:
            self.L.imba_integrator_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def preprocess(self):
        abi.check(self.L.imba_integrator_preprocess(self.h), "imba_integrator_preprocess")
        return self.L.imba_integrator_pixel_size(self.h)

    def render(self, iterations=1):
        abi.check(self.L.imba_integrator_render(self.h, iterations), "imba_integrator_render")

    def sync(self):
        abi.check(self.L.imba_integrator_sync(self.h), "imba_integrator_sync")

    def clear(self):
        abi.check(self.L.imba_integrator_clear(self.h), "imba_integrator_clear")

    def reset(self):
        """Integrator::reset (integrator.h:31, vcm.h:69-72): the progressive sequence (seeds, PPM / VCM radius schedule) starts over."""
        abi.check(self.L.imba_integrator_reset(self.h), "imba_integrator_reset")

    def set_framebuffer(self, tensor):
        assert tensor.is_cuda and tensor.dtype.is_floating_point and tensor.numel() == self.settings.width * self.settings.height * 3
        self._fb = tensor
        abi.check(self.L.imba_integrator_set_framebuffer(self.h, tensor.data_ptr()), "imba_integrator_set_framebuffer")

    def stream_ptr(self):
        return self.L.imba_integrator_stream(self.h)

    def download(self):
        img = np.empty((self.settings.height, self.settings.width, 3), np.float32)
        abi.check(self.L.imba_integrator_download(self.h, img.ctypes.data), "imba_integrator_download")
        return img

    def ray_counts(self):
        out = np.zeros(2, np.uint64)
        abi.check(self.L.imba_integrator_ray_counts(self.h, out.ctypes.data), "imba_integrator_ray_counts")
        return int(out[0]), int(out[1])

    def camera_rays(self, raster_xy):
        xy = np.ascontiguousarray(raster_xy, np.float32).reshape(-1, 2)
        rays = np.zeros(len(xy), abi.RAY)
        abi.check(self.L.imba_integrator_camera_rays(self.h, len(xy), xy.ctypes.data, rays.ctypes.data), "camera_rays")
        return rays

    def pt_shade_batch(self, rays, hits, states):
        n = len(rays)
        assert rays.dtype == abi.RAY and hits.dtype == abi.HIT and states.dtype == abi.PT_STATE
        rays_out = np.zeros(n, abi.RAY); states_out = np.zeros(n, abi.PT_STATE)
        sh_rays = np.zeros(n, abi.RAY); sh_states = np.zeros(n, abi.SHADOW_STATE)
        nc, ns = C.c_int(), C.c_int()
        abi.check(self.L.imba_pt_shade_batch(self.h, n, rays.ctypes.data, hits.ctypes.data, states.ctypes.data, rays_out.ctypes.data,
                                             states_out.ctypes.data, C.addressof(nc), sh_rays.ctypes.data, sh_states.ctypes.data, C.addressof(ns)),
                  "imba_pt_shade_batch")
        return rays_out[:nc.value], states_out[:nc.value], sh_rays[:ns.value], sh_states[:ns.value]

    def vcm_raygen_batch(self, light, iteration, base_radius, seed_base, count):
        rays = np.zeros(count, abi.RAY); states = np.zeros(count, abi.VCM_STATE)
        abi.check(self.L.imba_vcm_raygen_batch(self.h, int(light), iteration, base_radius, seed_base, count, rays.ctypes.data, states.ctypes.data), "imba_vcm_raygen_batch")
        return rays, states

    def vcm_stage_batch(self, stage, iteration, base_radius, rays, hits, states, cache_in=None):
        n = len(rays); slots = self.settings.num_connections + 1
        assert rays.dtype == abi.RAY and hits.dtype == abi.HIT and states.dtype == abi.VCM_STATE
        rays_out = np.zeros(n, abi.RAY); states_out = np.zeros(n, abi.VCM_STATE)
        sh_rays = np.zeros(n * slots, abi.RAY); sh_states = np.zeros(n * slots, abi.SHADOW_STATE)
        cache_out = np.zeros(2 * n + 16, abi.LIGHT_VERTEX)
        nc, ns, nv = C.c_int(), C.c_int(), C.c_int()
        cin = np.ascontiguousarray(cache_in) if cache_in is not None else np.zeros(0, abi.LIGHT_VERTEX)
        abi.check(self.L.imba_vcm_stage_batch(self.h, stage, iteration, base_radius, n, rays.ctypes.data, hits.ctypes.data, states.ctypes.data,
                                              cin.ctypes.data if len(cin) else None, len(cin), rays_out.ctypes.data, states_out.ctypes.data, C.addressof(nc),
                                              sh_rays.ctypes.data, sh_states.ctypes.data, C.addressof(ns), cache_out.ctypes.data, C.addressof(nv)), "imba_vcm_stage_batch")
        return rays_out[:nc.value], states_out[:nc.value], sh_rays[:ns.value], sh_states[:ns.value], cache_out[:nv.value]
