"""Host-buffer path only (bench.py's e2e leg) for tuning the staging pipeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imbatracer_b200 import api, abi
if os.environ.get('IMBA_LIB_VARIANT'):
    abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), os.environ['IMBA_LIB_VARIANT'])
from concurrent.futures import ThreadPoolExecutor
flat = bench.load_cornell(); scene = api.Scene(flat); n = bench.RAYS_PER_BATCH; L = abi.lib()
from tests import raygen
rays = [torch.from_numpy(np.ascontiguousarray(raygen.random_rays(flat, n, 5 + k)).view(np.uint8)).pin_memory() for k in range(4)]
hits = [torch.empty(n * 16, dtype=torch.uint8).pin_memory() for _ in range(4)]
pool = ThreadPoolExecutor(4)
def one(k): abi.check(L.imba_trace_host(scene.h, rays[k].data_ptr(), hits[k].data_ptr(), n, k & 1), "trace_host")
for threads in (1, 4):
    run = (lambda: [one(k) for k in range(4)]) if threads == 1 else (lambda: list(pool.map(one, range(4))))
    run(); t0 = time.perf_counter()
    for _ in range(5): run()
    dt = time.perf_counter() - t0
    print(os.environ.get('IMBA_LIB_VARIANT', 'base'), "threads", threads, "Mrays/s", round(4 * n * 5 / dt / 1e6, 1), "GB/s", round(4 * n * 5 * 48 / dt / 1e9, 1), flush=True)
# the ceiling beside it: pinned copies up and down at once, no kernels (bench.raw_pcie_bandwidth)
print("raw memcpy up+down GB/s", round(bench.raw_pcie_bandwidth(torch.cuda.synchronize), 1))
