"""Which occlusion rays of the bench's incoherent batch differ between the host-built and the device-built tree?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from imbatracer_b200 import api, abi
from tests import refapi
flat = bench.load_flat(); n = bench.RAYS_PER_BATCH
a = api.Scene(flat); b = api.Scene(flat, device_build=True)
batches = bench.make_batches(flat, n, 0x1234ABCD, scene=a)
for name in ("coherent_occluded", "incoherent_occluded"):
    rays = batches[name]
    d = api.rays_to_device(rays)
    ha = api.hits_from_device(a.trace(d, any_hit=True)); hb = api.hits_from_device(b.trace(d, any_hit=True))
    ca = api.hits_from_device(a.trace(d)); cb = api.hits_from_device(b.trace(d))
    torch.cuda.synchronize()
    diff = np.nonzero((ha["tri_id"] >= 0) != (hb["tri_id"] >= 0))[0]
    print(name, "any-hit flag differences:", len(diff), "closest-hit record differences:", int((ca.view(np.uint8).reshape(-1, 16) != cb.view(np.uint8).reshape(-1, 16)).any(1).sum()))
    sub = np.ascontiguousarray(rays[diff[:2000]])
    if len(sub):
        oa = refapi.trace_oracle(flat, sub, any_hit=True); oc = refapi.trace_oracle(flat, sub)
        print(" oracle any-hit says hit:", int((oa["tri_id"] >= 0).sum()), "of", len(sub), "; oracle closest says hit:", int((oc["tri_id"] >= 0).sum()))
        print(" host-tree any-hit == oracle:", int(((ha["tri_id"][diff[:2000]] >= 0) == (oa["tri_id"] >= 0)).sum()), " device-tree any-hit == oracle:", int(((hb["tri_id"][diff[:2000]] >= 0) == (oa["tri_id"] >= 0)).sum()))
        for i in range(min(6, len(sub))):
            k = diff[i]
            print(" ray", k, rays[k], "host any", ha[k], "dev any", hb[k], "host closest", ca[k], "dev closest", cb[k], "oracle closest", oc[i])
