"""A few progressive iterations of one integrator at 1080p on the bench scene (for launch lists / ncu):
   python tools/algo_iter.py lt [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imbatracer_b200 import api, abi
if os.environ.get('IMBA_LIB_VARIANT'):
    abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), os.environ['IMBA_LIB_VARIANT'])
algo = sys.argv[1]; iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flat = bench.load_flat(); scene = api.Scene(flat, device_build=True)
it = api.Integrator(scene, api.render_settings(flat, algo, 1920, 1080, seed=5))
it.preprocess(); it.render(1); it.sync()
t0 = time.perf_counter(); it.render(iters); it.sync(); dt = time.perf_counter() - t0
print(algo, "ms per iteration", dt / iters * 1e3, "rays", it.ray_counts())
