"""TEST INFRASTRUCTURE: drives the reference's own PathTracer::process_primary_rays (oracle/_ref) and the CUDA shading
stage (imba_pt_shade_batch) over the same injected wavefronts and reports how they differ."""
import os

import numpy as np

from imbatracer_b200 import abi, api
from tests import refapi


def initial_states(n_pixels, seed_base=0x1234ABCD, spp=1):
    L = abi.lib()
    st = np.zeros(n_pixels * spp, abi.PT_STATE)
    rng = np.zeros(n_pixels * spp, np.uint64)
    abi.check(L.imba_seed_states(seed_base, 0, len(rng), spp, rng.ctypes.data))
    st["pixel_id"] = np.arange(len(st)) // spp; st["sample_id"] = np.arange(len(st)) % spp
    st["rng"] = rng; st["throughput"] = 1.0
    return st


def compare_wave(ref_out, dev_out, tol=2e-4):
    """Returns a dict of mismatch statistics between (rays, states, shadow_rays, shadow_states[, fb]) tuples."""
    rr, rs, rsr, rss = ref_out[:4]
    dr, ds, dsr, dss = dev_out[:4]
    out = {"ref_cont": len(rr), "dev_cont": len(dr), "ref_shadow": len(rsr), "dev_shadow": len(dsr)}
    # continuation queue: match by pixel id (unique when spp == 1)
    rp, dp = {int(p): i for i, p in enumerate(rs["pixel_id"])}, {int(p): i for i, p in enumerate(ds["pixel_id"])}
    common = sorted(set(rp) & set(dp))
    out["cont_only_ref"] = len(set(rp) - set(dp)); out["cont_only_dev"] = len(set(dp) - set(rp))
    out["cont_same_order"] = bool(len(rr) == len(dr) and np.array_equal(rs["pixel_id"], ds["pixel_id"]))
    ri = np.array([rp[p] for p in common], int); di = np.array([dp[p] for p in common], int)
    if len(common):
        a, b = rr[ri], dr[di]
        def rel(x, y):
            return np.abs(x - y) / np.maximum(1e-3, np.maximum(np.abs(x), np.abs(y)))
        out["cont_rng_equal"] = float(np.mean(rs["rng"][ri] == ds["rng"][di]))
        out["cont_flags_equal"] = float(np.mean(rs["bounces_spec"][ri] == ds["bounces_spec"][di]))
        close = ((rel(a["org"], b["org"]).max(1) < tol) & (rel(a["dir"], b["dir"]).max(1) < tol) & (rel(a["tmin"], b["tmin"]) < tol)
                 & (rel(rs["throughput"][ri], ds["throughput"][di]).max(1) < tol) & (rel(rs["last_pdf"][ri], ds["last_pdf"][di]) < tol))
        out["cont_close"] = float(np.mean(close))
        out["cont_max_dir_err"] = float(np.abs(a["dir"] - b["dir"]).max())
    rp, dp = {int(p): i for i, p in enumerate(rss["pixel_id"])}, {int(p): i for i, p in enumerate(dss["pixel_id"])}
    common = sorted(set(rp) & set(dp))
    out["shadow_only_ref"] = len(set(rp) - set(dp)); out["shadow_only_dev"] = len(set(dp) - set(rp))
    if len(common):
        ri = np.array([rp[p] for p in common], int); di = np.array([dp[p] for p in common], int)
        a, b = rsr[ri], dsr[di]
        def rel(x, y):
            return np.abs(x - y) / np.maximum(1e-3, np.maximum(np.abs(x), np.abs(y)))
        close = ((rel(a["org"], b["org"]).max(1) < tol) & (rel(a["dir"], b["dir"]).max(1) < tol) & (rel(a["tmax"], b["tmax"]) < tol)
                 & (rel(a["tmin"], b["tmin"]) < tol) & (rel(rss["throughput"][ri], dss["throughput"][di]).max(1) < tol))
        out["shadow_close"] = float(np.mean(close))
    if len(ref_out) > 4 and len(dev_out) > 4:
        fa, fb = ref_out[4], dev_out[4]
        out["fb_ref_sum"] = float(fa.sum()); out["fb_dev_sum"] = float(fb.sum())
        out["fb_max_abs_err"] = float(np.abs(fa - fb).max())
    return out


def run_waves(scene_name, w=96, h=96, waves=4, max_path_len=10, layout="gpu", verbose=False, textured=False, envmap=False, phong=False):
    """Wave k feeds the REFERENCE's continuation queue of wave k-1 (traced by the oracle) to both implementations."""
    from tests.conftest import load_scene
    flat = load_scene(scene_name, layout)
    rs = refapi.RefScene(refapi.scene_path(scene_name), layout == "gpu")
    if envmap:    # environment map + EnvLight injected on the reference side
        assert rs.add_test_envmap() == 1
        flat = rs.flat(scene_name)
    if textured:  # textures, bump maps and texcoords injected on the reference side, exported to the device scene
        assert rs.add_test_textures() > 0
        flat = rs.flat(scene_name)
    if phong:     # materials that build the reference's Phong BxDF (brdfs.h:48-92), injected on the reference side
        assert rs.add_test_phong() > 0
        flat = rs.flat(scene_name)
        assert (flat.arrays["materials"]["kind"] == 4).any() and (flat.arrays["materials"]["kind"] == 5).any()   # Phong and OrenNayar
    scene = api.Scene(flat)
    integ = api.Integrator(scene, api.render_settings(flat, "pt", w, h, max_path_len=max_path_len))
    xy = np.stack(np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5), -1).reshape(-1, 2).astype(np.float32)
    rays = rs.camera_rays(w, h)
    cam_dev = integ.camera_rays(xy)
    results = [{"camera_max_dir_err": float(np.abs(cam_dev["dir"] - rays["dir"]).max()), "camera_org_equal": bool(np.array_equal(cam_dev["org"], rays["org"]))}]
    states = initial_states(w * h)
    for k in range(waves):
        hits = refapi.trace_oracle(flat, rays)
        integ.clear(); integ.sync()
        ref_out = rs.pt_shade_batch(w, h, rays, hits, states, max_path_len=max_path_len)
        dev = integ.pt_shade_batch(rays, hits, states)
        dev_out = dev + (integ.download(),)
        r = compare_wave(ref_out, dev_out); r["wave"] = k; r["in"] = len(rays)
        results.append(r)
        if verbose:
            print(scene_name, r, flush=True)
        rays, states = np.ascontiguousarray(ref_out[0]), np.ascontiguousarray(ref_out[1])
        if len(rays) == 0:
            break
    integ.close(); scene.close(); rs.close()
    return results
