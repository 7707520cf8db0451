import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import shading_harness as sh
for name in ["cornell_org", "cornell_specular_front", "cornell_water", "cornell_instances"]:
    try:
        res = sh.run_waves(name, verbose=True)
        print(name, res[0])
    except Exception as e:
        import traceback; traceback.print_exc()
