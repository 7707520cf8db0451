"""Generates the committed fixtures under tests/golden/ from the reference tree (run in the build container only):

  python tests/golden/make_fixtures.py [/root/reference [scene ...]]

  scene_<name>_<cpu|gpu>.npz  flattened scene exactly as the reference's loader + SBVH builder + adapters produce it
                              (frontend/build_scene.cpp:548, core/cpu_adapter.cpp, core/gpu_adapter.cpp, render/scene.cpp),
                              exported by oracle/_ref/libimba_ref.so (oracle/ref_driver.cpp: ref_scene_export)
  hits_<name>.npz             oracle closest-hit / occlusion results for the seeded ray batch of tests/raygen.py
  ref_<image>_128.npy         the reference's converged golden PNGs (test/references/*.png), box-filtered to 128x128 in
                              linear radiance (PNG = 8 bit, gamma 0.5, store_png.h:59-62)
  CornellBox-*.bvh            acceleration-structure caches written by the reference's own loader (load_bvh.cpp)
  scene_stilllife_proxy_*     the procedural Still Life stand-in (imbatracer_b200/procedural.py, detail refapi.PROXY_DETAIL) written as
                              OBJ/MTL/TGA by tests/obj_export.py and read back through the reference's loader + SBVH builder
  rng_kat.npz                 known-answer values of the reference's RNG (render/random.h:15-47) and of the primary-ray
                              seeding (ray_gen.h:57-72), produced by the unmodified reference headers
Also copies the reference's Cornell scenes to oracle/_ref/scenes (git-ignored scratch; the reference loader writes .bvh
caches beside the scene file, scene.cpp:67, so it must never run inside the reference tree)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import refapi, raygen  # noqa: E402

SCENES = ["cornell_org", "cornell_indirect", "cornell_water", "cornell_instances", "cornell_specular_front", "stilllife_proxy"]
IMAGES = ["cornell_org", "cornell_indirect", "cornell_specular", "cornell_specular_front", "cornell_water"]
RENDER_ONLY = ["cornell_specular"]   # scene arrays only (GPU layout): rendered against its golden image, no hit fixtures


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    scenes = refapi.ensure_ref_scenes(ref_root)
    only = sys.argv[2:]                      # optional: regenerate just these scenes
    for name in (only or SCENES):
        flat = {}
        for gpu in (False, True):
            rs = refapi.RefScene(refapi.scene_path(name), gpu)
            fs = rs.flat(name)
            fs.save(os.path.join(HERE, f"scene_{name}_{'gpu' if gpu else 'cpu'}.npz"))
            flat[gpu] = fs
            if not gpu:
                cam128 = rs.camera_rays(128, 128)
            rs.close()
        rays = raygen.test_batch(flat[False], cam128)
        out = {"camera_rays": cam128}
        for gpu in (False, True):
            tag = "gpu" if gpu else "cpu"
            closest = refapi.trace_oracle(flat[gpu], rays)
            occl = refapi.trace_oracle(flat[gpu], rays, any_hit=True)
            out[f"closest_{tag}_head"] = closest[:4096]
            out[f"closest_{tag}_sha"] = np.frombuffer(hashlib.sha256(closest.tobytes()).digest(), np.uint8)
            out[f"occluded_{tag}"] = np.packbits(occl["tri_id"] >= 0)
        np.savez_compressed(os.path.join(HERE, f"hits_{name}.npz"), **out)
        print(name, "tris", flat[False].tri_count, "rays", len(rays))

    if only:
        return
    for name in RENDER_ONLY:
        rs = refapi.RefScene(refapi.scene_path(name), True)
        rs.flat(name).save(os.path.join(HERE, f"scene_{name}_gpu.npz"))
        rs.close()
    # .bvh caches the reference's loader wrote next to the meshes while loading (render/scene.cpp:67, loaders/load_bvh.cpp)
    import shutil
    for f in ("CornellBox-Original.bvh", "CornellBox-Specular.bvh"):
        shutil.copy(os.path.join(scenes, "cornell", f), os.path.join(HERE, f))

    from PIL import Image
    for img in IMAGES:
        p = os.path.join(ref_root, "test", "references", f"ref_{img}.png")
        a = np.asarray(Image.open(p).convert("RGB"), np.float64) / 255.0
        lin = a ** 2.0  # stored value = radiance ** 0.5
        h, w, _ = lin.shape
        small = lin.reshape(128, h // 128, 128, w // 128, 3).mean(axis=(1, 3))
        np.save(os.path.join(HERE, f"ref_{img}_128.npy"), small.astype(np.float32))

    L = refapi.ref()
    vals = np.zeros(16, np.float32); st = np.zeros(1, np.uint64)
    L.ref_rng_floats(12345, 0, 16, vals.ctypes.data, st.ctypes.data)
    vals2 = np.zeros(16, np.float32); st2 = np.zeros(1, np.uint64)
    L.ref_rng_floats(0xDEADBEEFCAFE, 7, 16, vals2.ctypes.data, st2.ctypes.data)
    seeds = np.zeros(4096, np.uint64)
    L.ref_seed_states(0x1234ABCD, 0, 4096, 4, seeds.ctypes.data)
    sizes = np.zeros(16, np.int32); n = L.ref_struct_sizes(sizes.ctypes.data)
    np.savez(os.path.join(HERE, "rng_kat.npz"), floats_12345=vals, state_12345=st, floats_b=vals2, state_b=st2,
             seed_states=seeds, struct_sizes=sizes[:n])
    print("rng", vals[:3])


if __name__ == "__main__":
    main()
