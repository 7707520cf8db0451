"""One PT integrator at 1080p: preprocess + N iterations (for ncu launch lists / timing of the wavefront pipeline)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imbatracer_b200 import api, abi
if os.environ.get('IMBA_LIB_VARIANT'):
    abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), os.environ['IMBA_LIB_VARIANT'])
algo = sys.argv[1] if len(sys.argv) > 1 else "pt"; n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flat = bench.load_flat(); scene = api.Scene(flat)
it = api.Integrator(scene, api.render_settings(flat, algo, 1920, 1080, seed=5))
it.preprocess(); it.render(1); it.sync()
t0 = time.time(); it.render(n); it.sync(); dt = time.time() - t0
print(algo, "ms/iteration", dt / n * 1e3, "rays", it.ray_counts())
