"""TEST INFRASTRUCTURE: runs the reference's VCMIntegrator<algo>::process_light_rays / process_camera_rays (oracle/_ref) and
the CUDA stages (imba_vcm_stage_batch) over the same injected wavefronts -- light pass first (collecting the vertex cache),
then the camera pass with the REFERENCE's cache installed on both sides -- and reports how they differ."""
import os

import numpy as np

from imbatracer_b200 import abi, api
from tests import refapi

REF_ALGO = {"pt": 0, "bpt": 1, "vcm": 2, "ppm": 3, "vcm_pt": 4, "lt": 5}


def _rel(x, y):
    return np.abs(x - y) / np.maximum(1e-3, np.maximum(np.abs(x), np.abs(y)))


def _key(a):  # exact bit pattern of a float3 field, as a sortable key
    if len(a) == 0:
        return []
    return [bytes(x) for x in np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)]


def compare_states(rs, ds, rr, dr, tol):
    """Continuation queues matched by path / pixel id (unique at 1 spp)."""
    out = {"ref": len(rs), "dev": len(ds), "same_order": bool(len(rs) == len(ds) and np.array_equal(rs["pixel_id"], ds["pixel_id"]))}
    rp, dp = {int(p): i for i, p in enumerate(rs["pixel_id"])}, {int(p): i for i, p in enumerate(ds["pixel_id"])}
    common = sorted(set(rp) & set(dp)); out["only_one_side"] = len(set(rp) ^ set(dp))
    if common:
        ri, di = np.array([rp[p] for p in common]), np.array([dp[p] for p in common])
        out["rng_equal"] = float(np.mean(rs["rng"][ri] == ds["rng"][di])); out["len_equal"] = float(np.mean(rs["len_finite"][ri] == ds["len_finite"][di]))
        close = (_rel(rr["org"][ri], dr["org"][di]).max(1) < tol) & (_rel(rr["dir"][ri], dr["dir"][di]).max(1) < tol) & (_rel(rs["throughput"][ri], ds["throughput"][di]).max(1) < tol)
        for f in ("dVC", "dVCM", "dVM"):
            close &= _rel(rs[f][ri], ds[f][di]) < tol
        out["close"] = float(np.mean(close))
    return out


def compare_keyed(ra, da, key_field, fields, tol):
    """Records matched by the exact bits of `key_field` (positions are computed identically on both sides)."""
    rk, dk = {k: i for i, k in enumerate(_key(ra[key_field]))}, {k: i for i, k in enumerate(_key(da[key_field]))}
    common = sorted(set(rk) & set(dk))
    out = {"ref": len(ra), "dev": len(da), "only_one_side": len(set(rk) ^ set(dk))}
    if common:
        ri, di = np.array([rk[k] for k in common]), np.array([dk[k] for k in common])
        close = np.ones(len(common), bool)
        for f in fields:
            x, y = ra[f][ri], da[f][di]
            if x.dtype.kind == "f":
                e = _rel(x, y); close &= (e.max(1) if e.ndim > 1 else e) < tol
            else:
                close &= (x == y)
        out["close"] = float(np.mean(close))
    return out


def run(scene_name, algo, w=64, h=64, waves=4, iteration=3, layout="gpu", tol=5e-4, verbose=False, envmap=False, from_reference=False, phong=False):
    """from_reference: the scene arrays come from the reference's loader at run time (no committed fixture of that scene)."""
    from tests.conftest import load_scene
    rs = refapi.RefScene(refapi.scene_path(scene_name), layout == "gpu")
    flat = rs.flat(scene_name) if from_reference else load_scene(scene_name, layout)
    if envmap:
        assert rs.add_test_envmap() == 1
        flat = rs.flat(scene_name)
    if phong:   # materials that build the reference's Phong BxDF, injected on the reference side
        assert rs.add_test_phong() > 0
        flat = rs.flat(scene_name)
    scene = api.Scene(flat)
    n_light = w * h // 2
    integ = api.Integrator(scene, api.render_settings(flat, algo, w, h, light_path_count=n_light))
    base_radius = 2.0 * integ.preprocess()
    ra = REF_ALGO[algo]
    results = {"light": [], "camera": []}
    cache = np.zeros(0, abi.LIGHT_VERTEX)
    if algo != "vcm_pt":
        rays, states = integ.vcm_raygen_batch(True, iteration, base_radius, 0xABCDE, n_light)
        for k in range(waves):
            hits = refapi.trace_oracle(flat, rays)
            r_rays, r_states, r_sr, r_ss, r_fb, r_cache = rs.vcm_stage(ra, w, h, 0, iteration, base_radius, rays, hits, states, light_path_count=n_light)
            d_rays, d_states, d_sr, d_ss, d_cache = integ.vcm_stage_batch(0, iteration, base_radius, rays, hits, states)
            r = {"wave": k, "in": len(rays), "cont": compare_states(r_states, d_states, r_rays, d_rays, tol)}
            rsr = np.zeros(len(r_sr), [("org", "<f4", 3), ("dir", "<f4", 3), ("tmax", "<f4"), ("pixel", "<i4"), ("thr", "<f4", 3)])
            dsr = np.zeros(len(d_sr), rsr.dtype)
            for dst, a, b in ((rsr, r_sr, r_ss), (dsr, d_sr, d_ss)):
                dst["org"] = a["org"]; dst["dir"] = a["dir"]; dst["tmax"] = a["tmax"]; dst["pixel"] = b["pixel_id"]; dst["thr"] = b["throughput"]
            r["connect"] = compare_keyed(rsr, dsr, "org", ["dir", "tmax", "pixel", "thr"], tol)
            r["cache"] = compare_keyed(r_cache, d_cache, "pos", ["path_length", "mat", "out_dir", "normal", "geom_normal", "dVC", "dVCM", "dVM", "throughput"], tol)
            results["light"].append(r)
            if verbose:
                print(scene_name, algo, "light", r, flush=True)
            cache = np.concatenate([cache, r_cache])
            rays, states = np.ascontiguousarray(r_rays), np.ascontiguousarray(r_states)
            if len(rays) == 0:
                break
    if algo != "lt":
        rays, states = integ.vcm_raygen_batch(False, iteration, base_radius, 0x12345, w * h)
        for k in range(waves):
            hits = refapi.trace_oracle(flat, rays)
            integ.clear(); integ.sync()
            r_rays, r_states, r_sr, r_ss, r_fb, _ = rs.vcm_stage(ra, w, h, 1, iteration, base_radius, rays, hits, states, cache_in=cache, light_path_count=n_light)
            d_rays, d_states, d_sr, d_ss, _ = integ.vcm_stage_batch(1, iteration, base_radius, rays, hits, states, cache_in=cache)
            d_fb = integ.download()
            r = {"wave": k, "in": len(rays), "cache": len(cache), "cont": compare_states(r_states, d_states, r_rays, d_rays, tol)}
            rsr = np.zeros(len(r_sr), [("key", "<f4", 7), ("tmax", "<f4"), ("thr", "<f4", 3)])
            dsr = np.zeros(len(d_sr), rsr.dtype)
            for dst, a, b in ((rsr, r_sr, r_ss), (dsr, d_sr, d_ss)):
                dst["key"][:, :3] = a["org"]; dst["key"][:, 3:6] = a["dir"]; dst["key"][:, 6] = b["pixel_id"]; dst["tmax"] = a["tmax"]; dst["thr"] = b["throughput"]
            # shadow rays keyed by (origin, pixel) + rounded direction: NEE and connection rays of one vertex differ in direction
            for dst in (rsr, dsr):
                dst["key"][:, 3:6] = np.round(dst["key"][:, 3:6], 3)
            r["shadow"] = compare_keyed(rsr, dsr, "key", ["tmax", "thr"], tol)
            r["fb_ref_sum"] = float(r_fb.sum()); r["fb_dev_sum"] = float(d_fb.sum())
            r["fb_rel_err"] = float(np.abs(r_fb - d_fb).sum() / max(1e-9, np.abs(r_fb).sum()))
            results["camera"].append(r)
            if verbose:
                print(scene_name, algo, "camera", r, flush=True)
            rays, states = np.ascontiguousarray(r_rays), np.ascontiguousarray(r_states)
            if len(rays) == 0:
                break
    integ.close(); scene.close(); rs.close()
    return results
