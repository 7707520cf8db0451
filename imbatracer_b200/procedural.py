"""Procedural scenes: rigid instances of an existing mesh set on a jittered grid, and a large unique
terrain mesh.  Geometry, materials and lights are plain arrays; the acceleration structure is built by the product's own
host builder (imba_accel_build), not by the reference."""
import ctypes as C

import numpy as np

from . import abi
from .scene import FlatScene


def build_accel(flat: FlatScene, gpu=False):
    """Runs imba_accel_build (host, binned SAH) or imba_accel_build_gpu (device, linear BVH) on `flat` (meshes + instances) and
    stores the BVH2-layout arrays in it; flat.build_seconds = device time of the GPU mesh builds."""
    L = abi.lib()
    a = flat.arrays
    for k, dt in (("nodes", abi.NODE2), ("tris", np.float32), ("inst_nodes", abi.INSTANCE_NODE)):
        a.setdefault(k, np.zeros(0, dt))
    flat.accel_layout = 1
    d = flat.desc()
    secs = (C.c_float * 2)()
    h = L.imba_accel_build_gpu(C.byref(d), secs) if gpu else L.imba_accel_build(C.byref(d))
    if not h:
        raise RuntimeError("imba_accel_build failed: " + abi.last_error())
    flat.build_seconds = float(secs[0])
    p = [C.c_void_p() for _ in range(3)]; n = [C.c_int64() for _ in range(3)]; root = C.c_int32()
    abi.check(L.imba_accel_arrays(h, p[0], n[0], p[1], n[1], p[2], n[2], root))

    def grab(ptr, count, dt):
        dt = np.dtype(dt)
        buf = (C.c_uint8 * (count * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt, count=count).copy()

    a["nodes"] = grab(p[0], n[0].value, abi.NODE2)
    a["tris"] = grab(p[1], n[1].value * 4, np.float32)
    a["inst_nodes"] = grab(p[2], n[2].value, abi.INSTANCE_NODE)
    flat.root = root.value
    L.imba_accel_free(h)
    return flat


def _rigid(rng, pos, scale):
    ang = rng.random() * 2 * np.pi
    c, s = np.cos(ang) * scale, np.sin(ang) * scale
    m = np.array([[c, 0, s, pos[0]], [0, scale, 0, pos[1]], [-s, 0, c, pos[2]]], np.float64)
    r = m[:, :3]
    ri = np.linalg.inv(r)
    inv = np.concatenate([ri, (-ri @ m[:, 3])[:, None]], 1)
    return m.astype(np.float32), inv.astype(np.float32)


def instanced_grid(base: FlatScene, nx, ny, nz, spacing=3.0, seed=0x5EED, name="instanced"):
    """nx*ny*nz rigid instances (random rotation about y, jittered position, scale 0.8..1.2) of mesh 0 of `base`; the
    emissive triangles of the base scene are replicated per instance as triangle lights."""
    rng = np.random.default_rng(seed)
    s = FlatScene(); s.name = name; s.cam = base.cam.copy()
    a, b = s.arrays, base.arrays
    for k in ("mesh_vertex_base", "mesh_tri_base", "vertices", "normals", "texcoords", "geom_normals", "indices", "materials", "textures", "texels",
              "masks", "mask_buffer", "mask_indices", "mask_texcoords"):
        a[k] = b[k].copy()
    count = nx * ny * nz
    inst = np.zeros(count, abi.INSTANCE)
    base_lights = b["lights"]; tl = base_lights[base_lights["kind"] == 0]
    lights = np.zeros(count * len(tl), abi.LIGHT)
    i = 0
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    v = b["vertices"].reshape(-1, 4)[:, :3]
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                pos = (np.array([ix, iy, iz], np.float64) + (rng.random(3) - 0.5) * 0.3) * spacing
                m, inv = _rigid(rng, pos, 0.8 + 0.4 * rng.random())
                inst[i]["mesh_id"] = 0; inst[i]["mat"] = m.reshape(-1); inst[i]["inv_mat"] = inv.reshape(-1)
                for j, l in enumerate(tl):
                    o = lights[i * len(tl) + j]; o["kind"] = 0; o["intensity"] = l["intensity"]
                    for key in ("p0", "p1", "p2"):
                        o[key][:3] = m[:, :3] @ l[key][:3] + m[:, 3]
                w = v @ m[:, :3].T + m[:, 3]
                lo = np.minimum(lo, w.min(0)); hi = np.maximum(hi, w.max(0))
                i += 1
    a["instances"] = inst; a["lights"] = lights
    s.bsphere = np.array([*((lo + hi) / 2), np.linalg.norm(hi - lo) / 2], np.float32)
    c = (lo + hi) / 2
    s.cam[:3] = [c[0], c[1] + 0.1 * (hi[1] - lo[1]), hi[2] + 0.6 * (hi[2] - lo[2]) + 3.0]; s.cam[3:6] = [0, 0, -1]; s.cam[6:9] = [0, 1, 0]
    return build_accel(s)


def terrain(n=1024, seed=7, name="terrain"):
    """One unique mesh of 2*n*n triangles: a fractal height field over [-10, 10]^2 with four diffuse materials and two
    large triangle lights above it (stresses BVH size: ~128 bytes of nodes + triangles per triangle)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-10, 10, n + 1, dtype=np.float32) + np.float32(0.0137)  # keep grid lines off the coordinate planes
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    H = np.zeros_like(X)
    for o in range(1, 7):
        f = 0.15 * 2 ** o
        H += (np.sin(f * X + rng.random() * 6) * np.cos(f * Z + rng.random() * 6) / 2 ** o).astype(np.float32) * 2.5
    verts = np.stack([X, H, Z, np.ones_like(X)], -1).reshape(-1, 4).astype(np.float32)
    gy, gx = np.gradient(H, xs, xs)
    nrm = np.stack([-gx, np.ones_like(H), -gy], -1); nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    normals = np.concatenate([nrm.reshape(-1, 3), np.zeros((len(verts), 1))], 1).astype(np.float32)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    q = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[:-1, 1:], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 3)
    mat = ((q[:, 0] // (n + 1) // max(n // 8, 1) + q[:, 0] % (n + 1) // max(n // 8, 1)) % 4).astype(np.int32)
    # two emitter triangles high above the terrain
    lv = np.array([[-4, 9, -4, 1], [4, 9, -4, 1], [4, 9, 4, 1], [-4, 9, 4, 1]], np.float32)
    base = len(verts)
    verts = np.concatenate([verts, lv]); normals = np.concatenate([normals, np.tile(np.array([[0, -1, 0, 0]], np.float32), (4, 1))])
    q = np.concatenate([q, [[base, base + 1, base + 2], [base, base + 2, base + 3]]]); mat = np.concatenate([mat, [4, 5]]).astype(np.int32)
    indices = np.concatenate([q, mat[:, None]], 1).astype(np.int32)
    p0, p1, p2 = verts[q[:, 0], :3], verts[q[:, 1], :3], verts[q[:, 2], :3]
    gn = np.cross(p1 - p0, p2 - p0); area2 = np.linalg.norm(gn, axis=1, keepdims=True); gn = gn / np.maximum(area2, 1e-20)
    s = FlatScene(); s.name = name
    a = s.arrays
    a["mesh_vertex_base"] = np.array([0, len(verts)], np.int32); a["mesh_tri_base"] = np.array([0, len(q)], np.int32)
    a["vertices"] = verts.reshape(-1); a["normals"] = normals.reshape(-1); a["texcoords"] = np.zeros(len(verts) * 2, np.float32)
    a["geom_normals"] = np.concatenate([gn, np.zeros((len(q), 1))], 1).astype(np.float32).reshape(-1); a["indices"] = indices.reshape(-1)
    mats = np.zeros(6, abi.MATERIAL); mats["diffuse_tex"] = -1; mats["bump_tex"] = -1
    for k, col in enumerate([(0.7, 0.7, 0.7), (0.6, 0.3, 0.2), (0.2, 0.5, 0.25), (0.3, 0.35, 0.6)]):
        mats[k]["kd"] = col
    lights = np.zeros(2, abi.LIGHT)
    for k in range(2):
        t = len(q) - 2 + k
        mats[4 + k]["is_emitter"] = 1; mats[4 + k]["emit"] = (30, 30, 28); mats[4 + k]["emit_area"] = float(area2[t, 0] / 2)
        lights[k]["kind"] = 0; lights[k]["intensity"][:3] = (30, 30, 28)
        lights[k]["p0"][:3] = p0[t]; lights[k]["p1"][:3] = p1[t]; lights[k]["p2"][:3] = p2[t]
    a["materials"] = mats; a["lights"] = lights
    inst = np.zeros(1, abi.INSTANCE); eye = np.eye(3, 4, dtype=np.float32).reshape(-1); inst[0]["mat"] = eye; inst[0]["inv_mat"] = eye
    a["instances"] = inst
    a["textures"] = np.zeros(0, abi.TEXTURE); a["texels"] = np.zeros(0, np.float32)
    a["masks"] = np.zeros(0, abi.MASK); a["mask_buffer"] = np.zeros(0, np.uint8); a["mask_indices"] = np.zeros(0, np.int32); a["mask_texcoords"] = np.zeros(0, np.float32)
    lo, hi = verts[:, :3].min(0), verts[:, :3].max(0)
    s.bsphere = np.array([*((lo + hi) / 2), np.linalg.norm(hi - lo) / 2], np.float32)
    s.cam = np.array([0, 6, 14, 0, -0.35, -1, 0, 1, 0], np.float32)
    return build_accel(s)
