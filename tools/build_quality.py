"""Tree quality and build time of the three builders on the large scenes (run on the GPU box):
host binned SAH (accel_build.cpp) + host conversion, device LBVH + device collapse, device PLOC + device collapse."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imbatracer_b200 import api, procedural
from tests import raygen

n = 1 << 22
api.Scene(procedural.terrain(8), device_build=True).close()
for label, make in (("stilllife-proxy", lambda: procedural.still_life_proxy(1.0)), ("terrain_4.7M", lambda: procedural.terrain(1536))):
    t0 = time.perf_counter(); flat = make(); host_s = time.perf_counter() - t0
    rays = raygen.random_rays(flat, n, 11, tmax_mix=False); rays["tmin"] = 0
    d = api.rays_to_device(rays); hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    ref = None
    for builder in ("host_sah", "lbvh", "ploc"):
        os.environ["IMBA_GPU_BUILDER"] = builder
        t0 = time.perf_counter(); sc = api.Scene(flat, device_build=builder != "host_sah"); wall = time.perf_counter() - t0
        cnt = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
        sc.trace(d, hits, counters=cnt); torch.cuda.synchronize()
        h = hits.cpu().numpy().tobytes()
        ref = ref or h
        c = cnt.double().mean(0).cpu().numpy()
        for _ in range(3): sc.trace(d, hits)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): sc.trace(d, hits)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"scene": label, "builder": builder, "generate_and_host_build_s": round(host_s, 2) if builder == "host_sah" else None,
                          "scene_create_s": round(wall, 3), "device_ms": round(sc.build_info()["device_seconds"] * 1e3, 2), "levels": sc.build_info()["mesh_levels"],
                          "bytes": sc.stats()["bytes"], "nodes_per_ray": round(float(c[0]), 2), "tris_per_ray": round(float(c[1]), 2),
                          "closest_mrays_s": round(n * 5 / e0.elapsed_time(e1) / 1e3, 1), "same_hits": h == ref}))
        sc.close()
