"""Flattened scene container: numpy arrays behind an imba_scene_desc (include/imba_b200.h).

This is what a host program (the reference's `Scene`, render/scene.h) hands to imba_scene_create().
Arrays can be stored to / loaded from .npz so that tests and bench.py do not need the reference tree.
"""
import ctypes as C

import numpy as np

from . import abi

_ARRAYS = {
    # name: dtype
    "mesh_vertex_base": np.int32, "mesh_tri_base": np.int32, "vertices": np.float32, "normals": np.float32,
    "texcoords": np.float32, "geom_normals": np.float32, "indices": np.int32, "instances": abi.INSTANCE,
    "materials": abi.MATERIAL, "lights": abi.LIGHT, "textures": abi.TEXTURE, "texels": np.float32,
    "nodes": None, "tris": np.float32, "inst_nodes": abi.INSTANCE_NODE, "masks": abi.MASK, "mask_buffer": np.uint8,
    "mask_indices": np.int32, "mask_texcoords": np.float32,
}


class FlatScene:
    def __init__(self):
        self.arrays = {}
        self.accel_layout = 0          # 0 = MBVH4, 1 = BVH2
        self.root = 0
        self.env_tex = -1
        self.env_intensity = 0.0
        self.bsphere = np.zeros(4, np.float32)  # center xyz, radius
        self.cam = np.zeros(9, np.float32)      # pos, dir, up from the .scene file
        self.name = ""

    def __getattr__(self, k):
        arrays = object.__getattribute__(self, "arrays")
        if k in arrays:
            return arrays[k]
        raise AttributeError(k)

    @property
    def node_dtype(self):
        return abi.NODE4 if self.accel_layout == 0 else abi.NODE2

    @property
    def tri_count(self):
        return int(self.arrays["mesh_tri_base"][-1])

    def desc(self):
        """Builds the ctypes imba_scene_desc; keeps the arrays alive on the returned object."""
        a = self.arrays
        d = abi.SceneDesc()

        def ptr(name):
            arr = a.get(name)
            if arr is None or arr.size == 0:
                return None
            assert arr.flags["C_CONTIGUOUS"]
            return arr.ctypes.data

        d.mesh_count = len(a["mesh_tri_base"]) - 1
        for n in ("mesh_vertex_base", "mesh_tri_base", "vertices", "normals", "texcoords", "geom_normals", "indices",
                  "instances", "materials", "lights", "textures", "texels", "nodes", "tris", "inst_nodes", "masks",
                  "mask_buffer", "mask_indices", "mask_texcoords"):
            setattr(d, n, ptr(n))
        d.instance_count = len(a["instances"]); d.material_count = len(a["materials"]); d.light_count = len(a["lights"])
        d.texture_count = len(a["textures"]); d.texel_count = a["texels"].size // 4
        d.env_tex = self.env_tex; d.env_intensity = self.env_intensity
        d.bsphere_center[:] = [float(x) for x in self.bsphere[:3]]; d.bsphere_radius = float(self.bsphere[3])
        d.accel_layout = self.accel_layout; d.root = self.root; d.node_count = len(a["nodes"])
        d.tri_vec4_count = a["tris"].size // 4; d.inst_node_count = len(a["inst_nodes"])
        d.mask_count = len(a["masks"]); d.mask_buffer_size = a["mask_buffer"].size
        d.mask_texcoord_count = a["mask_texcoords"].size // 2
        d._keep = self
        return d

    def save(self, path):
        meta = np.array([self.accel_layout, self.root, self.env_tex], np.int64)
        np.savez_compressed(path, _meta=meta, _env_intensity=np.float32(self.env_intensity), _bsphere=self.bsphere, _cam=self.cam,
                            **{k: v.view(np.uint8) if v.dtype.fields else v for k, v in self.arrays.items()})

    @staticmethod
    def load(path, name=""):
        z = np.load(path)
        s = FlatScene()
        s.accel_layout, s.root, s.env_tex = (int(x) for x in z["_meta"])
        s.env_intensity = float(z["_env_intensity"]); s.bsphere = z["_bsphere"].astype(np.float32); s.cam = z["_cam"].astype(np.float32)
        s.name = name
        for k, dt in _ARRAYS.items():
            arr = np.ascontiguousarray(z[k])
            if k == "nodes":
                dt = s.node_dtype
            if np.dtype(dt).fields:
                arr = arr.view(dt)
            s.arrays[k] = arr
        return s


def from_desc(d, accel_layout):
    """Copies everything an imba_scene_desc points at into a FlatScene (used on descriptors exported by a host)."""
    s = FlatScene()

    def grab(ptr, count, dtype):
        dtype = np.dtype(dtype)
        if not ptr or count == 0:
            return np.zeros(0, dtype)
        buf = (C.c_uint8 * (count * dtype.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).copy()

    mc = d.mesh_count
    a = s.arrays
    a["mesh_vertex_base"] = grab(d.mesh_vertex_base, mc + 1, np.int32)
    a["mesh_tri_base"] = grab(d.mesh_tri_base, mc + 1, np.int32)
    nv, nt = int(a["mesh_vertex_base"][-1]), int(a["mesh_tri_base"][-1])
    a["vertices"] = grab(d.vertices, nv * 4, np.float32); a["normals"] = grab(d.normals, nv * 4, np.float32)
    a["texcoords"] = grab(d.texcoords, nv * 2, np.float32); a["geom_normals"] = grab(d.geom_normals, nt * 4, np.float32)
    a["indices"] = grab(d.indices, nt * 4, np.int32)
    a["instances"] = grab(d.instances, d.instance_count, abi.INSTANCE)
    a["materials"] = grab(d.materials, d.material_count, abi.MATERIAL)
    a["lights"] = grab(d.lights, d.light_count, abi.LIGHT)
    a["textures"] = grab(d.textures, d.texture_count, abi.TEXTURE)
    a["texels"] = grab(d.texels, d.texel_count * 4, np.float32)
    s.accel_layout = accel_layout; s.root = d.root
    a["nodes"] = grab(d.nodes, d.node_count, s.node_dtype)
    a["tris"] = grab(d.tris, d.tri_vec4_count * 4, np.float32)
    a["inst_nodes"] = grab(d.inst_nodes, d.inst_node_count, abi.INSTANCE_NODE)
    a["masks"] = grab(d.masks, d.mask_count, abi.MASK)
    a["mask_buffer"] = grab(d.mask_buffer, d.mask_buffer_size, np.uint8)
    a["mask_indices"] = grab(d.mask_indices, nt * 4, np.int32)
    a["mask_texcoords"] = grab(d.mask_texcoords, d.mask_texcoord_count * 2, np.float32)
    s.env_tex = d.env_tex; s.env_intensity = d.env_intensity
    s.bsphere = np.array([d.bsphere_center[0], d.bsphere_center[1], d.bsphere_center[2], d.bsphere_radius], np.float32)
    return s
