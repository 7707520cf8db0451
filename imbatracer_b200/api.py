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

    def __init__(self, flat: FlatScene):
        self.L = abi.lib()
        self.flat = flat
        d = flat.desc()
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
