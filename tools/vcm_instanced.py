import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imbatracer_b200 import api, procedural
base = bench.load_cornell()
flat = procedural.instanced_grid(base, 14, 14, 14)
scene = api.Scene(flat)
for algo in sys.argv[1:] or ["pt", "lt", "bpt", "ppm", "vcm"]:
    it = api.Integrator(scene, api.render_settings(flat, algo, 1280, 720, seed=3))
    t0 = time.time(); it.preprocess(); tp = time.time() - t0
    it.render(1); it.sync()
    t0 = time.time(); it.render(2); it.sync(); dt = (time.time() - t0) / 2
    print(algo, "preprocess", round(tp, 3), "s; ms/iteration", round(dt * 1e3, 2), "rays", it.ray_counts(), flush=True)
    it.close()
