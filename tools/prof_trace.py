"""Launch sequence for `ncu --set full` captures of the traversal kernel on the bench workload (or an extra scene):
16 generation launches (make_batches through imba_trace_host: 2 x 8 chunks) or none for the extra scenes, then one warm-up
and one measured launch of each batch, in BATCH_ORDER.  usage: tools/prof_trace.py [stilllife|terrain|instanced]
  ncu ... -k regex:trace_ -s 20 -c 4 python tools/prof_trace.py            (stilllife: skip 16 + 4 warm-ups)
  ncu ... -k regex:trace_ -s 2 -c 2 python tools/prof_trace.py terrain     (closest + occluded on 4 Mi random rays)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imbatracer_b200 import api, abi, procedural
if os.environ.get('IMBA_LIB_VARIANT'):
    abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), os.environ['IMBA_LIB_VARIANT'])
which = sys.argv[1] if len(sys.argv) > 1 else "stilllife"
n = bench.RAYS_PER_BATCH
if which == "stilllife":
    flat = bench.load_flat(); scene = api.Scene(flat)
    batches = bench.make_batches(flat, n, 0x1234ABCD, scene=scene)
    order = bench.BATCH_ORDER
else:
    from tests import raygen
    base = bench.load_cornell("cornell_water")
    flat = procedural.terrain(1536) if which == "terrain" else procedural.instanced_grid(base, 14, 14, 14)
    scene = api.Scene(flat)
    rays = raygen.random_rays(flat, n, 11, tmax_mix=False); rays["tmin"] = 0
    batches = {"closest": rays, "occluded": rays}
    order = [("closest", 0), ("occluded", 1)]
dev = {k: api.rays_to_device(v) for k, v in batches.items()}
hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
for rep in range(2):
    for name, any_hit in order:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); scene.trace(dev[name], hits, any_hit=bool(any_hit), count=n); e1.record(); torch.cuda.synchronize()
        if rep: print(which, name, round(e0.elapsed_time(e1), 3), "ms")
