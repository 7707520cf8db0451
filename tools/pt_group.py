"""PT 1080p samples/s as a function of iterations_per_pass (1 spp per iteration)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imbatracer_b200 import api
flat = bench.load_flat(); scene = api.Scene(flat)
for g in (1, 2, 4, 8, 16):
    it = api.Integrator(scene, api.render_settings(flat, "pt", 1920, 1080, spp=1, seed=5, iterations_per_pass=g))
    it.preprocess(); it.render(g); it.sync()
    n = max(16, 2 * g)
    t0 = time.perf_counter(); it.render(n); it.sync(); dt = time.perf_counter() - t0
    print("iterations per pass", g, "ms per iteration", round(dt / n * 1e3, 3), "M samples/s", round(1920 * 1080 * n / dt / 1e6, 1), flush=True)
    it.close()
