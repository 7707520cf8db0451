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


# ---- "stilllife-proxy": a stand-in for the reference's Still Life benchmark scene ---------------------------------------------
# test/scenes/stilllife/still_life.{scene,mtl} are in the reference tree but still_life.obj is not.  The proxy keeps the
# camera of still_life.scene and the 23 materials of still_life.mtl (restated below as data, in file order) and replaces the
# missing geometry by procedurally tessellated objects of the same kinds (draped cloth, fruit, vase with flowers, metal bowl,
# glass bottle, glasses of coloured water, gems, ring, mirror, three area lights).  It is reported as "stilllife-proxy", never
# as Still Life.
STILL_LIFE_CAMERA = (4.00539, 4.33851, 4.25406, -0.937134, -0.229336, -0.26303, -0.200025, 0.970634, -0.13364123)
STILL_LIFE_MTL = [
    ("Cap_1SG", dict(Ns=0.0, Ka=(0.59, 0.59, 0.59), illum=6)),
    ("initialShadingGroup", dict(Ns=96.078431, Kd=(0.4, 0.4, 0.4), Ks=(0.5, 0.5, 0.5), illum=6)),
    ("apple_001", dict(Ns=15.0, map_Kd="apple_001.tga", illum=2)),
    ("apple_002", dict(Ns=15.0, map_Kd="apple_002.tga", illum=2)),
    ("bottle_cap", dict(Kd=(1.0, 1.0, 1.0), illum=2)),
    ("cloth_bottom", dict(map_Kd="fabric_bot.tga", illum=2)),
    ("cloth_top", dict(map_Kd="fabric_top.tga", illum=2)),
    ("leaf", dict(Kd=(0.2, 0.6, 0.01), illum=2)),
    ("petals", dict(Kd=(0.8, 0.1, 0.01), illum=2)),
    ("vase", dict(Kd=(1.0, 1.0, 1.0), illum=2)),
    ("walls", dict(Kd=(0.273831, 0.273831, 0.273831), illum=2)),
    ("bowl", dict(Ns=96.078431, Ks=(1.0, 1.0, 1.0), illum=2)),
    ("cap_metal", dict(Ns=96.078431, Kd=(0.2, 0.2, 0.2), Ks=(0.5, 0.5, 0.5), illum=2)),
    ("mirror", dict(Ns=96.078431, Ks=(1.0, 1.0, 1.0), illum=5)),
    ("ring", dict(Ns=96.078431, Kd=(0.2, 0.2, 0.2), Ks=(0.5, 0.5, 0.5), illum=2)),
    ("diamond", dict(Ks=(1.0, 1.0, 1.0), Tf=(1.0, 1.0, 1.0), Ni=3.0, illum=7)),
    ("glass", dict(Ks=(1.0, 1.0, 1.0), Tf=(1.0, 1.0, 1.0), Ni=1.7, illum=7)),
    ("water_blue", dict(Ks=(0.2, 0.6, 0.9), Tf=(0.2, 0.6, 0.9), Ni=1.33, illum=7)),
    ("water_clear", dict(Ks=(1.0, 1.0, 1.0), Tf=(1.0, 1.0, 1.0), Ni=1.33, illum=7)),
    ("water_green", dict(Ks=(0.3, 1.0, 0.3), Tf=(0.3, 1.0, 0.3), Ni=1.33, illum=7)),
    ("back_light", dict(Ke=(100.0, 100.0, 100.0), illum=2)),
    ("fill_light", dict(Ke=(500.0, 500.0, 500.0), illum=2)),
    ("key_light", dict(Ke=(1000.0, 1000.0, 1000.0), illum=2)),
]


def mtl_to_materials(mtl):
    """Material records + texture file list for an MTL table, by the rules of the reference's convert_materials
    (frontend/build_scene.cpp:55-180): index 0 is the dummy diffuse material; illum 5 = mirror(eta 1, kappa Ns, scale Ks),
    illum 7 = glass(Ni, Tf, Ks), Ks > 0 = glossy, else diffuse; map_Kd becomes the diffuse texture; Ke > 0 marks an emitter
    (one duplicated material per emissive triangle is appended later, build_scene.cpp:221-241)."""
    mats = np.zeros(len(mtl) + 1, abi.MATERIAL); mats["diffuse_tex"] = -1; mats["bump_tex"] = -1
    mats[0]["kd"] = (1.0, 1.0, 1.0)            # DiffuseMaterial() default colour (materials.h)
    tex_names, emissive = [], {}
    for i, (_, m) in enumerate(mtl, start=1):
        o = mats[i]
        kd, ks, ke = m.get("Kd", (0, 0, 0)), m.get("Ks", (0, 0, 0)), m.get("Ke", (0, 0, 0))
        if max(ke) > 0:
            emissive[i] = ke
        tex = -1
        if m["illum"] == 5:
            o["kind"] = 1; o["eta"] = 1.0; o["kappa"] = m.get("Ns", 0.0); o["ks"] = ks
            continue
        if m["illum"] == 7:
            o["kind"] = 2; o["eta"] = m.get("Ni", 0.0); o["ks"] = ks; o["tf"] = m.get("Tf", (0, 0, 0))
            continue
        if "map_Kd" in m:
            if m["map_Kd"] not in tex_names:
                tex_names.append(m["map_Kd"])
            tex = tex_names.index(m["map_Kd"])
        if max(ks) > 0:
            o["kind"] = 3; o["exponent"] = m.get("Ns", 0.0); o["ks"] = ks; o["diffuse_tex"] = tex
            o["kd"] = (0, 0, 0) if tex >= 0 else kd   # GlossyMaterial's sampler constructor leaves diffuse_color_ at its default
        else:
            o["kind"] = 0; o["diffuse_tex"] = tex
            if tex < 0:
                o["kd"] = kd
    return mats, tex_names, emissive


def _proxy_texture(name, size=256):
    """Seeded 8-bit RGB pattern per texture file name: mottled fruit skins, woven fabrics."""
    rng = np.random.default_rng(abs(hash_name(name)) % (2 ** 32))
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    if name.startswith("apple"):
        base = np.array([0.75, 0.12, 0.08] if name.endswith("001.tga") else [0.55, 0.7, 0.15], np.float32)
        n = sum(np.sin(2 * np.pi * (fx * x + fy * y) + ph) / (1 + k) for k, (fx, fy, ph) in enumerate(rng.integers(1, 9, (6, 3)).astype(np.float32)))
        img = base[None, None, :] * (0.8 + 0.12 * n[..., None]) + 0.08 * rng.random((size, size, 1), dtype=np.float32)
    else:
        warp = (np.floor(x * 32) + np.floor(y * 32)) % 2
        weave = 0.5 + 0.5 * np.sin(2 * np.pi * 64 * x) * np.sin(2 * np.pi * 64 * y)
        c0, c1 = (np.array([0.7, 0.65, 0.5]), np.array([0.35, 0.1, 0.1])) if "top" in name else (np.array([0.25, 0.3, 0.5]), np.array([0.8, 0.8, 0.75]))
        img = (c0[None, None] * warp[..., None] + c1[None, None] * (1 - warp[..., None])) * (0.75 + 0.25 * weave[..., None])
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


def hash_name(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xffffffff
    return h


class _MeshBuilder:
    def __init__(self):
        self.P, self.N, self.T, self.F, self.M, self.nv = [], [], [], [], [], 0

    def add(self, P, N, T, F, mat):
        self.P.append(P.astype(np.float32)); self.N.append(N.astype(np.float32)); self.T.append(T.astype(np.float32))
        self.F.append(F.astype(np.int64) + self.nv); self.M.append(np.full(len(F), mat, np.int32)); self.nv += len(P)

    def param(self, fn, nu, nv, mat, flip=False, uv_scale=(1.0, 1.0), faceted=False, urange=(0.0, 1.0), vrange=(0.0, 1.0)):
        """Tessellates P = fn(u, v) on an nu x nv grid (two triangles per cell); normals from central differences of fn."""
        nu, nv = max(int(nu), 3), max(int(nv), 2)
        u = np.linspace(urange[0], urange[1], nu + 1); v = np.linspace(vrange[0], vrange[1], nv + 1)
        U, V = np.meshgrid(u, v, indexing="ij")
        P = fn(U, V)
        e = 1e-4
        n = np.cross(fn(U + e, V) - fn(U - e, V), fn(U, V + e) - fn(U, V - e))
        n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-30)
        if flip:
            n = -n
        idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
        a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
        F = np.stack([a, b, c, a, c, d], -1).reshape(-1, 3)
        if flip:
            F = F[:, ::-1]
        P = P.reshape(-1, 3); n = n.reshape(-1, 3)
        T = np.stack([U * uv_scale[0], V * uv_scale[1]], -1).reshape(-1, 2)
        if faceted:
            P = P[F].reshape(-1, 3); T = T[F].reshape(-1, 2)
            g = np.cross(P[1::3] - P[0::3], P[2::3] - P[0::3]); g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-30)
            n = np.repeat(g, 3, axis=0); F = np.arange(len(P)).reshape(-1, 3)
        self.add(P, n, T, F, mat)

    def box(self, lo, hi, mat, inward=False):
        lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        for ax in range(3):
            for side in (0, 1):
                a1, a2 = (ax + 1) % 3, (ax + 2) % 3
                c = np.zeros((4, 3)); c[:, ax] = hi[ax] if side else lo[ax]
                c[:, a1] = [lo[a1], hi[a1], hi[a1], lo[a1]]; c[:, a2] = [lo[a2], lo[a2], hi[a2], hi[a2]]
                nrm = np.zeros(3); nrm[ax] = 1.0 if side else -1.0
                F = np.array([[0, 1, 2], [0, 2, 3]])
                if not side:
                    F = F[:, ::-1]
                if inward:
                    nrm = -nrm; F = F[:, ::-1]
                self.add(c, np.tile(nrm, (4, 1)), np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64), F, mat)

    def quad(self, p0, eu, ev, mat, n=1):
        p0, eu, ev = (np.asarray(x, np.float64) for x in (p0, eu, ev))
        self.param(lambda U, V: p0 + U[..., None] * eu + V[..., None] * ev, max(n, 1), max(n, 1), mat) if n >= 3 else self._quad1(p0, eu, ev, mat, n)

    def _quad1(self, p0, eu, ev, mat, n):
        nrm = np.cross(eu, ev); nrm /= np.linalg.norm(nrm)
        for i in range(n):
            for j in range(n):
                o = p0 + eu * (i / n) + ev * (j / n)
                c = np.array([o, o + eu / n, o + eu / n + ev / n, o + ev / n])
                self.add(c, np.tile(nrm, (4, 1)), np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64), np.array([[0, 1, 2], [0, 2, 3]]), mat)


def _revolve(profile, cx, y0, cz):
    """Surface of revolution about the vertical axis through (cx, cz): profile = [(radius, height), ...] polyline,
    parametrised by arc length along v, angle along u."""
    pr = np.asarray(profile, np.float64)
    seg = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(pr, axis=0), axis=1))]); seg /= seg[-1]

    def fn(U, V):
        Vc = np.clip(V, 0.0, 1.0)
        r = np.interp(Vc, seg, pr[:, 0]); h = np.interp(Vc, seg, pr[:, 1])
        ang = 2 * np.pi * U
        return np.stack([cx + r * np.cos(ang), y0 + h, cz + r * np.sin(ang)], -1)
    return fn


def still_life_proxy(detail=1.0, name="stilllife-proxy", accel="host"):
    """~1.4 M triangles at detail 1 (tessellation densities scale with sqrt(detail)); one mesh, one identity instance,
    24 + (emissive triangle count) materials, four 256x256 textures.  accel: "host" (binned SAH), "gpu" (LBVH) or None."""
    k = float(np.sqrt(detail))
    mid = {n: i + 1 for i, (n, _) in enumerate(STILL_LIFE_MTL)}
    b = _MeshBuilder()
    ty = 3.0                                   # table top height; the camera of still_life.scene looks at about (0, 3.3, 3.1)
    # room and table
    b.box((-4.0, 0.0, -1.9), (6.0, 8.0, 8.1), mid["walls"], inward=True)
    b.box((-1.6, ty - 0.2, 1.5), (1.6, ty, 4.7), mid["initialShadingGroup"])
    for lx in (-1.5, 1.3):
        for lz in (1.6, 4.4):
            b.box((lx, 0.0, lz), (lx + 0.2, ty - 0.2, lz + 0.2), mid["initialShadingGroup"])

    def cloth(extent, lift, hang, phase):
        def fn(U, V):
            x = -extent + 2 * extent * U; z = 3.1 - extent + 2 * extent * V
            d = np.maximum(np.maximum(np.abs(x) - 1.6, np.abs(z - 3.1) - 1.6), 0.0)
            drop = hang * (1.0 - np.exp(-6.0 * d)) + 0.35 * d
            y = ty + lift + 0.004 * np.sin(41.0 * x + phase) * np.sin(37.0 * z) * np.exp(-8.0 * d) - drop
            push = 0.05 * np.sin(23.0 * (x + z) + phase) * np.minimum(d * 6.0, 1.0)
            return np.stack([x + push * np.sign(x) * (np.abs(x) > 1.6), y, z + push * np.sign(z - 3.1) * (np.abs(z - 3.1) > 1.6)], -1)
        return fn
    b.param(cloth(2.0, 0.004, 0.9, 0.3), 300 * k, 300 * k, mid["cloth_bottom"], flip=True, uv_scale=(6, 6))
    b.param(cloth(1.85, 0.012, 0.45, 1.7), 500 * k, 500 * k, mid["cloth_top"], flip=True, uv_scale=(8, 8))

    def blob(c, r, squash, wob):
        def fn(U, V):
            th = np.pi * V; ph = 2 * np.pi * U
            rr = r * (1.0 + wob * np.sin(3 * ph) * np.sin(th) ** 2 - 0.12 * np.cos(th) ** 4)
            return np.stack([c[0] + rr * np.sin(th) * np.cos(ph), c[1] + squash * rr * np.cos(th), c[2] + rr * np.sin(th) * np.sin(ph)], -1)
        return fn
    pole = (0.004, 0.996)
    b.param(blob((0.30, ty + 0.33, 3.40), 0.16, 0.92, 0.03), 200 * k, 200 * k, mid["apple_001"], flip=True, vrange=pole)
    b.param(blob((0.95, ty + 0.17, 2.45), 0.15, 0.95, 0.02), 200 * k, 200 * k, mid["apple_002"], flip=True, vrange=pole)
    # vase with flowers
    vase = [(0.004, 0.02), (0.16, 0.02), (0.22, 0.25), (0.12, 0.6), (0.07, 0.85), (0.1, 1.0), (0.085, 1.0), (0.055, 0.85), (0.1, 0.6), (0.2, 0.25), (0.14, 0.05), (0.004, 0.05)]
    b.param(_revolve(vase, -0.6, ty, 2.6), 300 * k, 300 * k, mid["vase"])
    rng = np.random.default_rng(11)
    for f in range(9):
        ang, lean, h = rng.random() * 2 * np.pi, 0.15 + 0.35 * rng.random(), 0.5 + 0.5 * rng.random()
        top = np.array([-0.6 + lean * np.cos(ang), ty + 1.0 + h, 2.6 + lean * np.sin(ang)])
        base = np.array([-0.6, ty + 0.9, 2.6])

        def stem(U, V, base=base, top=top):
            p = base[None, None] + (top - base)[None, None] * V[..., None] + np.array([0, 0.08, 0])[None, None] * np.sin(np.pi * V)[..., None]
            a = 2 * np.pi * U
            return p + 0.006 * np.stack([np.cos(a), np.zeros_like(a), np.sin(a)], -1)
        b.param(stem, 8, 24 * k, mid["leaf"])
        for l in range(7):
            t = 0.2 + 0.1 * l; o = base + (top - base) * t; la = rng.random() * 2 * np.pi

            def leaf(U, V, o=o, la=la):
                w = 0.05 * np.sin(np.pi * V) * (U - 0.5) * 2; ln = 0.22 * V
                x = ln * np.cos(la) - w * np.sin(la); z = ln * np.sin(la) + w * np.cos(la)
                return np.stack([o[0] + x, o[1] + 0.08 * np.sin(2.5 * V) - 0.15 * (U - 0.5) ** 2, o[2] + z], -1)
            b.param(leaf, 16 * k, 16 * k, mid["leaf"])
        for p in range(12):
            pa = 2 * np.pi * p / 12 + 0.2 * f

            def petal(U, V, top=top, pa=pa):
                w = 0.035 * np.sin(np.pi * V) * (U - 0.5) * 2; ln = 0.13 * V
                x = ln * np.cos(pa) - w * np.sin(pa); z = ln * np.sin(pa) + w * np.cos(pa)
                return np.stack([top[0] + x, top[1] + 0.05 * V ** 2 + 0.02 * np.cos(np.pi * (U - 0.5)), top[2] + z], -1)
            b.param(petal, 12 * k, 12 * k, mid["petals"])
    # metal bowl (outer and inner wall in one profile)
    bowl = [(0.004, 0.0), (0.2, 0.0), (0.4, 0.12), (0.46, 0.3), (0.44, 0.3), (0.38, 0.13), (0.19, 0.025), (0.004, 0.025)]
    b.param(_revolve(bowl, 0.3, ty + 0.012, 3.4), 256 * k, 256 * k, mid["bowl"])
    # glass bottle, its water, cap parts
    bottle = [(0.004, 0.0), (0.13, 0.0), (0.14, 0.05), (0.14, 0.5), (0.05, 0.72), (0.045, 0.9), (0.035, 0.9), (0.04, 0.72), (0.128, 0.5), (0.128, 0.06), (0.004, 0.02)]
    b.param(_revolve(bottle, -0.9, ty + 0.012, 3.8), 256 * k, 256 * k, mid["glass"])
    b.param(_revolve([(0.004, 0.021), (0.127, 0.061), (0.127, 0.4), (0.004, 0.4)], -0.9, ty + 0.012, 3.8), 128 * k, 64 * k, mid["water_clear"])
    b.param(_revolve([(0.047, 0.86), (0.05, 0.86), (0.05, 0.93), (0.004, 0.93)], -0.9, ty + 0.012, 3.8), 64 * k, 16 * k, mid["bottle_cap"])
    b.param(_revolve([(0.05, 0.84), (0.052, 0.84), (0.052, 0.87), (0.05, 0.87)], -0.9, ty + 0.012, 3.8), 64 * k, 8 * k, mid["cap_metal"])
    b.param(_revolve([(0.004, 0.931), (0.04, 0.931)], -0.9, ty + 0.012, 3.8), 32 * k, 2, mid["Cap_1SG"], flip=True)
    # three glasses with coloured water
    tumbler = [(0.004, 0.0), (0.09, 0.0), (0.11, 0.32), (0.102, 0.32), (0.084, 0.02), (0.004, 0.02)]
    for (gx, gz), water in zip(((0.55, 4.25), (0.95, 3.95), (1.25, 4.35)), ("water_blue", "water_clear", "water_green")):
        b.param(_revolve(tumbler, gx, ty + 0.012, gz), 160 * k, 96 * k, mid["glass"])
        b.param(_revolve([(0.004, 0.021), (0.083, 0.021), (0.097, 0.24), (0.004, 0.24)], gx, ty + 0.012, gz), 128 * k, 48 * k, mid[water])
    # ring and gems
    def torus(U, V):
        a, c = 2 * np.pi * U, 2 * np.pi * V
        return np.stack([1.1 + (0.09 + 0.012 * np.cos(c)) * np.cos(a), ty + 0.025 + 0.012 * np.sin(c), 3.0 + (0.09 + 0.012 * np.cos(c)) * np.sin(a)], -1)
    b.param(torus, 192 * k, 48 * k, mid["ring"], flip=True)
    gem = [(0.003, 0.0), (0.05, 0.055), (0.035, 0.075), (0.003, 0.075)]
    for gx, gz in ((1.1, 3.0), (1.3, 3.25), (0.8, 2.9), (1.35, 2.8), (0.65, 3.1)):
        b.param(_revolve(gem, gx, ty + 0.02, gz), 12, 3, mid["diamond"], faceted=True)
    # mirror on the far wall with a frame, facing +x
    b.quad((-3.98, 2.4, 1.4), (0, 0, 3.4), (0, 3.2, 0), mid["mirror"])
    b.box((-3.995, 2.3, 1.3), (-3.985, 5.7, 4.9), mid["ring"])
    # area lights (their triangles become the scene's triangle lights)
    b.quad((1.6, 7.2, 4.6), (0.5, 0, 0), (0, -0.1, 0.5), mid["key_light"], n=2)
    b.quad((3.0, 5.5, -0.5), (0.8, 0, 0.3), (-0.2, 0.8, 0), mid["fill_light"], n=2)
    b.quad((-3.5, 6.5, 6.5), (1.0, 0, 0), (0, 0.3, 1.0), mid["back_light"], n=2)

    P = np.concatenate(b.P); N = np.concatenate(b.N); T = np.concatenate(b.T); F = np.concatenate(b.F); M = np.concatenate(b.M)
    p0, p1, p2 = P[F[:, 0]], P[F[:, 1]], P[F[:, 2]]
    gn = np.cross(p1 - p0, p2 - p0); alen = np.linalg.norm(gn, axis=1)
    keep = alen > 1e-14                                                      # drop degenerate cells of the parametrisations
    F, M, gn, alen = F[keep], M[keep], gn[keep], alen[keep]
    # geometric normals in float32 exactly as build_scene.cpp:298-302 computes them: normalize(cross(p1 - p0, p2 - p0)),
    # normalize(a) = a * (1 / length(a)) (core/float3.h:74-78,122-128)
    f32 = np.float32
    ea, eb = (P[F[:, 1]] - P[F[:, 0]]).astype(f32), (P[F[:, 2]] - P[F[:, 0]]).astype(f32)
    cr = np.stack([ea[:, 1] * eb[:, 2] - ea[:, 2] * eb[:, 1], ea[:, 2] * eb[:, 0] - ea[:, 0] * eb[:, 2], ea[:, 0] * eb[:, 1] - ea[:, 1] * eb[:, 0]], -1)
    ln = np.sqrt(cr[:, 0] * cr[:, 0] + cr[:, 1] * cr[:, 1] + cr[:, 2] * cr[:, 2])
    gn = cr * (f32(1.0) / ln)[:, None]
    mats, tex_names, emissive = mtl_to_materials(STILL_LIFE_MTL)
    # one duplicated material + one triangle light per emissive triangle, in triangle order (build_scene.cpp:221-241, 616-626)
    em_tris = np.nonzero(np.isin(M, list(emissive)))[0]
    obj_material = M.copy()                                                  # material per face as an OBJ names it
    extra = np.zeros(len(em_tris), abi.MATERIAL); lights = np.zeros(len(em_tris), abi.LIGHT)
    for j, t in enumerate(em_tris):
        extra[j] = mats[M[t]]; extra[j]["is_emitter"] = 1; extra[j]["emit"] = emissive[int(M[t])]; extra[j]["emit_area"] = np.float32(alen[t] * 0.5)
        lights[j]["kind"] = 0; lights[j]["intensity"][:3] = emissive[int(M[t])]
        lights[j]["p0"][:3] = P[F[t, 0]]; lights[j]["p1"][:3] = P[F[t, 1]]; lights[j]["p2"][:3] = P[F[t, 2]]
        M[t] = len(mats) + j
    mats = np.concatenate([mats, extra])
    tex_bytes = [_proxy_texture(n) for n in tex_names]
    textures = np.zeros(len(tex_bytes), abi.TEXTURE); texels = []; off = 0
    for i, tb in enumerate(tex_bytes):
        h, w, _ = tb.shape
        textures[i]["width"] = w; textures[i]["height"] = h; textures[i]["offset"] = off; off += w * h
        rgba = np.concatenate([tb.astype(np.float32) / np.float32(255.0), np.ones((h, w, 1), np.float32)], -1)
        texels.append(rgba.reshape(-1))
    s = FlatScene(); s.name = name
    a = s.arrays
    nvert, ntri = len(P), len(F)
    a["mesh_vertex_base"] = np.array([0, nvert], np.int32); a["mesh_tri_base"] = np.array([0, ntri], np.int32)
    a["vertices"] = np.concatenate([P, np.ones((nvert, 1), np.float32)], 1).astype(np.float32).reshape(-1)
    a["normals"] = np.concatenate([N, np.zeros((nvert, 1), np.float32)], 1).astype(np.float32).reshape(-1)
    a["texcoords"] = T.astype(np.float32).reshape(-1)
    a["geom_normals"] = np.concatenate([gn, np.zeros((ntri, 1), np.float32)], 1).reshape(-1)
    a["indices"] = np.concatenate([F, M[:, None]], 1).astype(np.int32).reshape(-1)
    a["materials"] = mats; a["lights"] = lights; a["textures"] = textures
    a["texels"] = np.concatenate(texels) if texels else np.zeros(0, np.float32)
    inst = np.zeros(1, abi.INSTANCE); eye = np.eye(3, 4, dtype=np.float32).reshape(-1); inst[0]["mat"] = eye; inst[0]["inv_mat"] = eye
    a["instances"] = inst
    a["masks"] = np.zeros(0, abi.MASK); a["mask_buffer"] = np.zeros(0, np.uint8); a["mask_indices"] = np.zeros(0, np.int32); a["mask_texcoords"] = np.zeros(0, np.float32)
    lo, hi = P.min(0), P.max(0)
    s.bsphere = np.array([*((lo + hi) / 2), np.linalg.norm(hi - lo) / 2], np.float32)
    s.cam = np.array(STILL_LIFE_CAMERA, np.float32)
    s.mtl, s.tex_names, s.tex_bytes, s.obj_material = STILL_LIFE_MTL, tex_names, tex_bytes, obj_material
    if accel is None:
        return s
    return build_accel(s, gpu=(accel == "gpu"))
