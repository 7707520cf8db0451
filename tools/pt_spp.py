"""PT 1080p samples/s as a function of the samples per pixel per wavefront pass (the reference's --spp)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imbatracer_b200 import api
flat = bench.load_flat(); scene = api.Scene(flat)
for spp in (1, 2, 4, 8):
    it = api.Integrator(scene, api.render_settings(flat, "pt", 1920, 1080, spp=spp, seed=5))
    it.preprocess(); it.render(1); it.sync()
    n = max(1, 8 // spp)
    t0 = time.perf_counter(); it.render(n); it.sync(); dt = time.perf_counter() - t0
    print("spp", spp, "ms per iteration", round(dt / n * 1e3, 3), "M samples/s", round(1920 * 1080 * spp * n / dt / 1e6, 1), flush=True)
    it.close()
