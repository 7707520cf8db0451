"""Device-resident traversal timing only (the four batches of bench.py), for kernel tuning runs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imbatracer_b200 import api, abi
import hashlib
if os.environ.get('IMBA_LIB_VARIANT'):
    abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), os.environ['IMBA_LIB_VARIANT'])
flat = bench.load_flat(); scene = api.Scene(flat, device_build=bool(os.environ.get("IMBA_DEVICE_BUILD"))); n = bench.RAYS_PER_BATCH
cache = "/tmp/bench_batches.npz"
if os.path.exists(cache):
    z = np.load(cache); batches = {k: z[k].view(api.abi.RAY).reshape(-1) for k in z.files}
else:
    batches = bench.make_batches(flat, n, 0x1234ABCD, scene=scene)
    np.savez(cache, **{k: v.view(np.uint8) for k, v in batches.items()})
dev = {k: api.rays_to_device(v) for k, v in batches.items()}
hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
res = {}
for name, any_hit in bench.BATCH_ORDER:
    for _ in range(3): scene.trace(dev[name], hits, any_hit=bool(any_hit), count=n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): scene.trace(dev[name], hits, any_hit=bool(any_hit), count=n)
    e1.record(); torch.cuda.synchronize()
    res[name] = round(n * 10 / e0.elapsed_time(e1) / 1e3, 1)
    h = hits.cpu().numpy()
    cnt = torch.zeros((n, 4), dtype=torch.int32, device='cuda'); scene.trace(dev[name], hits, any_hit=bool(any_hit), count=n, counters=cnt)
    c = cnt.double().mean(0).cpu().numpy(); res[name + '_sha'] = [round(float(c[0]), 2), round(float(c[1]), 2), int(cnt[:, 3].max())]
    res_sha = None
    res[name + '_sha'] = res[name + '_sha'] + [hashlib.sha256((h[:, 0] >= 0).tobytes() if any_hit else h.tobytes()).hexdigest()[:8]]
print({k: v for k, v in os.environ.items() if k.startswith("IMBA_")}, json.dumps(res), "Mrays/s total", round(4 * n / sum(n / v / 1e3 for k, v in res.items() if not k.endswith('_sha')) / 1e3, 1))
