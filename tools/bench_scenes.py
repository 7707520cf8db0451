"""Traversal throughput + oracle spot check on the procedural scenes (instanced ~20 M triangles, unique terrain)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imbatracer_b200 import api, abi, procedural
from imbatracer_b200.scene import FlatScene
from tests import refapi, raygen

def run(flat, n=1 << 22, label=""):
    t0 = time.time(); scene = api.Scene(flat); t_up = time.time() - t0
    st = scene.stats()
    # incoherent: random rays inside the bounding sphere; coherent: camera rays
    rays = raygen.random_rays(flat, n, 11, tmax_mix=False); rays["tmin"] = 0
    d = api.rays_to_device(rays); hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    scene.trace(d, hits, counters=cnt); torch.cuda.synchronize()
    c = cnt.double().mean(0).cpu().numpy()
    out = {"scene": label, "triangles": int(flat.tri_count), "instances": int(len(flat.instances)), "bvh8": st, "upload_s": round(t_up, 2)}
    for any_hit in (0, 1):
        for _ in range(3): scene.trace(d, hits, any_hit=bool(any_hit))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): scene.trace(d, hits, any_hit=bool(any_hit))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out["occluded" if any_hit else "closest"] = {"ms": round(ms, 3), "mrays_s": round(n / ms / 1e3, 1)}
    bpr = 48 + 80 * c[0] + 48 * c[1] + 64 * c[2]
    out["per_ray"] = {"nodes": round(c[0], 2), "tris": round(c[1], 2), "instances": round(c[2], 2), "bytes": round(bpr, 1)}
    out["algorithmic_GBs"] = round(n * bpr / (out["closest"]["ms"] * 1e-3) / 1e9, 1)
    h = api.hits_from_device(scene.trace(d, hits))[: 1 << 14]
    want = refapi.trace_oracle(flat, np.ascontiguousarray(rays[: 1 << 14]))
    out["oracle_sample_bit_exact"] = bool(h.tobytes() == want.tobytes()); out["hit_rate"] = round(float((h["tri_id"] >= 0).mean()), 3)
    scene.close()
    print(json.dumps(out), flush=True)
    return out

which = sys.argv[1:] or ["instanced", "terrain"]
base = FlatScene.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scene_cornell_water_gpu.npz"))
if "instanced" in which:
    t0 = time.time(); s = procedural.instanced_grid(base, 14, 14, 14); print("build instanced", round(time.time() - t0, 1), "s", flush=True)
    run(s, label="cornell_water x 2744 instances")
if "terrain" in which:
    t0 = time.time(); s = procedural.terrain(int(os.environ.get("TERRAIN_N", "1536"))); print("build terrain", round(time.time() - t0, 1), "s", flush=True)
    run(s, label="terrain (host binned SAH)")
    t0 = time.time(); procedural.build_accel(s, gpu=True); print("gpu lbvh: device", round(s.build_seconds * 1e3, 2), "ms, whole call", round(time.time() - t0, 2), "s", flush=True)
    run(s, label="terrain (GPU LBVH)")
