"""Traversal of the bench batches on three structures of the same mesh: the product's host builder (binned SAH, what the bench uses),
the tree the REFERENCE's SplitBvhBuilder builds (bench.reference_built_flat) and the device builder (PLOC).  Same rays, same hits."""
import os, sys, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imbatracer_b200 import api
n = bench.RAYS_PER_BATCH
flat = bench.load_flat(); scene = api.Scene(flat)
batches = bench.make_batches(flat, n, 0x1234ABCD, scene=scene)
dev = {k: api.rays_to_device(v) for k, v in batches.items()}
hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
def run(sc, label):
    res = {}
    for name, any_hit in bench.BATCH_ORDER:
        for _ in range(3): sc.trace(dev[name], hits, any_hit=bool(any_hit), count=n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): sc.trace(dev[name], hits, any_hit=bool(any_hit), count=n)
        e1.record(); torch.cuda.synchronize()
        h = hits.cpu().numpy()
        cnt = torch.zeros((n, 4), dtype=torch.int32, device='cuda'); sc.trace(dev[name], hits, any_hit=bool(any_hit), count=n, counters=cnt)
        c = cnt.double().mean(0).cpu().numpy()
        res[name] = [round(n * 10 / e0.elapsed_time(e1) / 1e3, 1), round(float(c[0]), 2), round(float(c[1]), 2),
                     hashlib.sha256((h[:, 0] >= 0).tobytes() if any_hit else h.tobytes()).hexdigest()[:8]]
    print(label, json.dumps(res), "bytes", sc.stats()["bytes"], flush=True)
run(scene, "product host builder (binned SAH):")
ref_flat, how = bench.reference_built_flat()
if ref_flat is not None:
    s2 = api.Scene(ref_flat); run(s2, f"reference SBVH ({how}):"); s2.close()
s3 = api.Scene(flat, device_build=True); run(s3, "device builder (PLOC):"); s3.close()
