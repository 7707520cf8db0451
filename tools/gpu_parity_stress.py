"""One-off parity stress beyond the unit tests: N seeds x 4 Mi random rays per scene, CUDA path vs the CPU oracle, closest hit
bit-exact and occlusion flag-equal (coplanar hit-or-miss class excepted and counted).  usage: tools/gpu_parity_stress.py [seeds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imbatracer_b200 import api, abi, procedural
if os.environ.get('IMBA_LIB_VARIANT'):
    abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), os.environ['IMBA_LIB_VARIANT'])
from tests import raygen, refapi

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = 1 << 22
base = bench.load_cornell("cornell_water")
scenes = {"stilllife-proxy": bench.load_flat(), "cornell_water": base,
          "instanced_2744": procedural.instanced_grid(base, 14, 14, 14), "terrain_4.7M": procedural.terrain(1536)}
total = bad_closest = bad_occ = 0
for name, flat in scenes.items():
    sc = api.Scene(flat)
    for seed in range(int(os.environ.get("SEED0", "100")), int(os.environ.get("SEED0", "100")) + seeds):
        rays = raygen.random_rays(flat, n, seed)
        d = api.rays_to_device(rays)
        got = api.hits_from_device(sc.trace(d)); occ = api.hits_from_device(sc.trace(d, any_hit=True)); torch.cuda.synchronize()
        t0 = time.time()
        want = refapi.trace_oracle(flat, rays); wocc = refapi.trace_oracle(flat, rays, any_hit=True)
        g = got.view(np.uint32).reshape(-1, 4); w = want.view(np.uint32).reshape(-1, 4)
        bc = int((g != w).any(axis=1).sum()); bo = int(((occ["tri_id"] >= 0) != (wocc["tri_id"] >= 0)).sum())
        total += n; bad_closest += bc; bad_occ += bo
        if bc or bo:   # keep the offending rays for analysis (tests/test_bench_parity_gpu.py documents the known exception classes)
            ic = np.nonzero((g != w).any(axis=1))[0]; io = np.nonzero((occ["tri_id"] >= 0) != (wocc["tri_id"] >= 0))[0]
            os.makedirs("gpurun_out", exist_ok=True)
            np.savez("gpurun_out/stress_%s_%d.npz" % (name.replace(".", "_"), seed), closest_rays=rays[ic], closest_got=got[ic], closest_want=want[ic],
                     occ_rays=rays[io], occ_got=occ[io], occ_want=wocc[io])
            for i in ic: print("  closest", i, rays[i], "got", got[i], "want", want[i])
            for i in io: print("  occluded", i, rays[i], "got", occ[i], "want", wocc[i])
        print(name, "seed", seed, "closest mismatches", bc, "occlusion flag mismatches", bo, "oracle s", round(time.time() - t0, 1), flush=True)
    sc.close()
print("rays per mode", total, "closest mismatches", bad_closest, "occlusion mismatches", bad_occ)
