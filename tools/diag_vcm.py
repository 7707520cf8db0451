import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from imbatracer_b200 import api, abi
from tests import refapi
from tests.conftest import load_scene
REF_ALGO = {"pt": 0, "bpt": 1, "vcm": 2, "ppm": 3, "vcm_pt": 5, "lt": 4}  # UserSettings::Algorithm order: PT, BPT, VCM, PPM, LT, VCM_PT? checked below
def grmse(a, b): return float(np.sqrt(np.mean((np.clip(a,0,1)**0.5 - np.clip(b,0,1)**0.5)**2)))
scenes = sys.argv[1:] or ["cornell_org", "cornell_water"]
W = H = 128; N = 64
for name in scenes:
    flat = load_scene(name, "gpu")
    sc = api.Scene(flat)
    rs = refapi.RefScene(os.path.join(refapi.REF_SCENES, "cornell", name + ".scene"), True)
    imgs = {}
    for algo in ["pt", "vcm_pt", "lt", "bpt", "ppm", "vcm"]:
        try:
            it = api.Integrator(sc, api.render_settings(flat, algo, W, H, seed=3))
            px = it.preprocess()
            t0 = time.time(); it.render(N); img = it.download() / N; dt = time.time() - t0
            prim, shad = it.ray_counts()
            it.close()
            imgs[algo] = img
            line = f"{name} {algo:7s} dev mean {img.mean():.5f} nan {int(np.isnan(img).sum())} max {img.max():.2f} rays {prim} {shad} {dt*1e3/N:.2f} ms/iter px {px:.5f}"
            ra = {"pt": 0, "bpt": 1, "vcm": 2, "ppm": 3, "vcm_pt": 4, "lt": 5}[algo]
            rimg, st = rs.render(ra, W, H, iterations=N if algo not in ("vcm", "ppm") else 32)
            rimg = rimg / (N if algo not in ("vcm", "ppm") else 32)
            line += f" | ref mean {rimg.mean():.5f} rays {st['primary_rays']} {st['shadow_rays']} rmse(dev,ref) {grmse(img, rimg):.4f}"
            if "pt" in imgs: line += f" rmse(dev,devpt) {grmse(img, imgs['pt']):.4f}"
            print(line, flush=True)
        except Exception as e:
            import traceback; traceback.print_exc()
    sc.close(); rs.close()
