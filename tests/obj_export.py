"""Test infrastructure: writes a procedurally built FlatScene (imbatracer_b200/procedural.py) as the files the reference's
own loader reads -- Wavefront OBJ + MTL (loaders/load_obj.cpp), uncompressed 24-bit TGA textures (loaders/load_tga.cpp) and a
`.scene` file (frontend/build_scene.cpp:316-503) -- so that the same scene can be loaded by the reference (SBVH builder,
adapters, integrators) and compared with the arrays the generator made."""
import os
import struct

import numpy as np


def write_tga(path, rgb8):
    """rgb8[y, x] is image row y as the reference's Image stores it; the file holds rows bottom-up in BGR (load_tga.cpp:55-66)."""
    h, w, _ = rgb8.shape
    with open(path, "wb") as f:
        f.write(bytes([0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0]) + struct.pack("<HHBB", w, h, 24, 0))
        f.write(np.ascontiguousarray(rgb8[::-1, :, ::-1]).tobytes())


def write_scene(flat, directory, base="stilllife_proxy"):
    os.makedirs(directory, exist_ok=True)
    a = flat.arrays
    P = a["vertices"].reshape(-1, 4)[:, :3]; N = a["normals"].reshape(-1, 4)[:, :3]; T = a["texcoords"].reshape(-1, 2)
    F = a["indices"].reshape(-1, 4)[:, :3]
    names = ["<dummy>"] + [n for n, _ in flat.mtl]
    with open(os.path.join(directory, base + ".mtl"), "w") as f:
        for n, m in flat.mtl:
            f.write(f"newmtl {n}\n")
            for k in ("Ns", "Ni"):
                if k in m:
                    f.write(f"{k} {float(np.float32(m[k])):.9g}\n")
            for k in ("Ka", "Kd", "Ks", "Ke", "Tf"):
                if k in m:
                    f.write(f"{k} " + " ".join(f"{float(np.float32(x)):.9g}" for x in m[k]) + "\n")
            if "map_Kd" in m:
                f.write(f"map_Kd {m['map_Kd']}\n")
            f.write(f"illum {m['illum']}\n\n")
    for n, tb in zip(flat.tex_names, flat.tex_bytes):
        write_tga(os.path.join(directory, n), tb)
    with open(os.path.join(directory, base + ".obj"), "w") as f:
        f.write(f"mtllib {base}.mtl\n")
        for n in names[1:]:
            f.write(f"usemtl {n}\n")          # registers the materials in table order (load_obj.cpp:199-211)
        f.write("\n".join("v %.9g %.9g %.9g" % tuple(p) for p in P.tolist()) + "\n")
        f.write("\n".join("vt %.9g %.9g" % tuple(t) for t in T.tolist()) + "\n")
        f.write("\n".join("vn %.9g %.9g %.9g" % tuple(n) for n in N.tolist()) + "\n")
        cur = -1
        lines = []
        for (i0, i1, i2), m in zip((F + 1).tolist(), flat.obj_material.tolist()):
            if m != cur:
                lines.append(f"usemtl {names[m]}"); cur = m
            lines.append(f"f {i0}/{i0}/{i0} {i1}/{i1}/{i1} {i2}/{i2}/{i2}")
        f.write("\n".join(lines) + "\n")
    c = flat.cam
    with open(os.path.join(directory, base + ".scene"), "w") as f:
        f.write(f"mesh {base}.obj\naccel {base}.bvh\n\npos {c[0]:.9g} {c[1]:.9g} {c[2]:.9g}\ndir {c[3]:.9g} {c[4]:.9g} {c[5]:.9g}\nup {c[6]:.9g} {c[7]:.9g} {c[8]:.9g}\n")
    return os.path.join(directory, base + ".scene")
