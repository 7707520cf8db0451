"""A small end-to-end run (seconds): host-built and device-built structure, traversal, host-buffer path, PT (grouped), VCM / PPM / BPT
(pipelined, tail, merging), LT -- prints hit counts and mean radiances, which must agree between the two structures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imbatracer_b200 import api, procedural
from tests import raygen
flat = procedural.still_life_proxy(0.01)
for device_build in (False, True):
    sc = api.Scene(flat, device_build=device_build)
    rays = raygen.random_rays(flat, 20000, 3)
    h = api.hits_from_device(sc.trace(api.rays_to_device(rays)))
    o = api.hits_from_device(sc.trace(api.rays_to_device(rays), any_hit=True))
    hh = sc.trace_host(np.ascontiguousarray(rays))
    torch.cuda.synchronize()
    print("device_build", device_build, "hits", int((h["tri_id"] >= 0).sum()), int((o["tri_id"] >= 0).sum()), int((hh["tri_id"] >= 0).sum()))
    for algo, kw in (("pt", dict(iterations_per_pass=3)), ("vcm", {}), ("ppm", {}), ("bpt", {}), ("lt", {})):
        it = api.Integrator(sc, api.render_settings(flat, algo, 48, 32, seed=2, **kw))
        it.preprocess(); it.render(4)
        img = it.download(); it.close()
        print(" ", algo, float(img.mean()))
    sc.close()
