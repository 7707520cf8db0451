import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import vcm_harness as vh
for scene in ["cornell_org", "cornell_specular_front"]:
    for algo in ["vcm", "bpt", "ppm", "lt", "vcm_pt"]:
        try:
            vh.run(scene, algo, waves=3, verbose=True)
        except Exception:
            import traceback; traceback.print_exc()
