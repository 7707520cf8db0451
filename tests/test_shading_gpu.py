"""GPU parity of the shading stage and of the device path tracer.
 * wavefront level: the CUDA stage (sort by material -> shade -> compact) against the REFERENCE'S OWN
   PathTracer::process_primary_rays (pt.cpp:82-160, compiled from /root/reference into oracle/_ref) on the same injected
   rays / hits / 40-byte states: same survivors in the same order, identical RNG states (i.e. the same number and order
   of random draws), identical bounce/specular flags, rays / throughputs / pdfs within 2e-4 relative (the differences are
   the device's sinf/cosf/powf roundings), identical emitter contributions;
 * image level: converged device renders against the reference's integrator on the oracle and against the reference's
   golden PNG (gamma-0.5 RMSE, the metric of test/run_bench.py:113-129)."""
import os

import numpy as np
import pytest

from imbatracer_b200 import abi
from tests import refapi
from tests.conftest import GOLDEN, load_scene

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refapi.have_ref() or not os.path.isdir(refapi.REF_SCENES), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("scene", ["cornell_org", "cornell_specular_front", "cornell_water", "cornell_instances", "stilllife_proxy"])
def test_pt_shading_stage_matches_reference_wavefronts(scene):
    from tests import shading_harness as sh
    res = sh.run_waves(scene, w=96, h=96, waves=5)
    assert res[0]["camera_org_equal"] and res[0]["camera_max_dir_err"] < 5e-6   # closed-form camera vs matrix inversion
    for r in res[1:]:
        assert r["ref_cont"] == r["dev_cont"] and r["ref_shadow"] == r["dev_shadow"], r
        assert r["cont_only_ref"] == 0 and r["cont_only_dev"] == 0 and r["shadow_only_ref"] == 0 and r["shadow_only_dev"] == 0, r
        if r["ref_cont"]:
            assert r["cont_same_order"] and r["cont_rng_equal"] == 1.0 and r["cont_flags_equal"] == 1.0, r
            assert r["cont_close"] >= 0.999, r
        if r["ref_shadow"]:
            assert r["shadow_close"] >= 0.999, r
        assert r["fb_max_abs_err"] <= 1e-4 * max(1.0, r["fb_ref_sum"]), r


@needs_ref
@pytest.mark.parametrize("scene", ["cornell_org", "cornell_specular_front"])
def test_textured_and_bump_mapped_shading_matches_reference(scene):
    """TextureSampler::sample (texture_sampler.h:18-46), Material::bump (materials.h:34-55) and the textured get_bsdf
    branches: procedural textures / bump maps / texcoords are injected into the reference's scene, exported, and both
    sides shade the same wavefronts."""
    from tests import shading_harness as sh
    res = sh.run_waves(scene, w=96, h=96, waves=4, textured=True)
    for r in res[1:]:
        assert r["ref_cont"] == r["dev_cont"] and r["ref_shadow"] == r["dev_shadow"], r
        if r["ref_cont"]:
            assert r["cont_same_order"] and r["cont_rng_equal"] == 1.0 and r["cont_flags_equal"] == 1.0, r
            assert r["cont_close"] >= 0.995, r
        if r["ref_shadow"]:
            assert r["shadow_close"] >= 0.995, r


@needs_ref
@pytest.mark.parametrize("scene", ["cornell_org", "cornell_specular_front", "cornell_water"])
def test_phong_brdf_shading_matches_reference(scene):
    """The Phong BxDF (render/materials/brdfs.h:48-92; named by the north star) and OrenNayar (brdfs.h:94-134), both instantiated by no
    material of the reference's loader: materials that build BSDF(isect, Phong, nullptr) / BSDF(isect, OrenNayar, nullptr) replace the
    diffuse / glossy materials alternately on the reference side (oracle/ref_driver.cpp: PhongTestMaterial, OrenNayarTestMaterial), the
    scene is exported with kinds IMBA_MAT_PHONG / IMBA_MAT_OREN_NAYAR, and both sides shade the same PT
    wavefronts -- Phong::eval in the NEE, Phong::sample (power-cosine lobe about the mirrored direction, random draws in g++'s
    argument order) and Phong::pdf in the MIS weights: same survivors in the same order, identical RNG states."""
    from tests import shading_harness as sh
    res = sh.run_waves(scene, w=96, h=96, waves=4, phong=True)
    for r in res[1:]:
        assert r["ref_cont"] == r["dev_cont"] and r["ref_shadow"] == r["dev_shadow"], r
        assert r["cont_only_ref"] == 0 and r["cont_only_dev"] == 0 and r["shadow_only_ref"] == 0 and r["shadow_only_dev"] == 0, r
        if r["ref_cont"]:
            assert r["cont_same_order"] and r["cont_rng_equal"] == 1.0 and r["cont_flags_equal"] == 1.0, r
            assert r["cont_close"] >= 0.998, r
        if r["ref_shadow"]:
            assert r["shadow_close"] >= 0.998, r
        assert r["fb_max_abs_err"] <= 1e-4 * max(1.0, r["fb_ref_sum"]), r


@needs_ref
def test_environment_map_shading_matches_reference():
    """EnvMap::radiance on escaped rays (pt.cpp:95-116) and EnvLight::sample_direct as one of the NEE lights
    (light.h:443-453, 516-530): a procedural environment map is added to the reference's scene and exported."""
    from tests import shading_harness as sh
    res = sh.run_waves("cornell_org", w=96, h=96, waves=4, envmap=True)
    for r in res[1:]:
        assert r["ref_cont"] == r["dev_cont"] and r["ref_shadow"] == r["dev_shadow"], r
        if r["ref_cont"]:
            assert r["cont_same_order"] and r["cont_rng_equal"] == 1.0 and r["cont_flags_equal"] == 1.0 and r["cont_close"] >= 0.995, r
        if r["ref_shadow"]:
            assert r["shadow_close"] >= 0.995, r
        assert r["fb_ref_sum"] > 0 and r["fb_max_abs_err"] <= 2e-4 * max(1.0, r["fb_ref_sum"]), r
    # (wave 0 sees only the inside of the box; rays escape through the open front from the first bounce on)
    assert sum(r["fb_ref_sum"] for r in res[2:]) > 0


def gamma_rmse(a, b):
    return float(np.sqrt(np.mean((np.clip(a, 0, 1) ** 0.5 - np.clip(b, 0, 1) ** 0.5) ** 2)))


@pytest.mark.parametrize("scene,tol_ref", [("cornell_org", 0.05), ("cornell_water", 0.16), ("stilllife_proxy", 0.2)])
def test_device_path_tracer_converges_to_reference(scene, tol_ref):
    from imbatracer_b200 import api
    flat = load_scene(scene, "gpu")
    sc = api.Scene(flat)
    spp_iters = 256
    integ = api.Integrator(sc, api.render_settings(flat, "pt", 128, 128, seed=7))
    px = integ.preprocess()
    integ.render(spp_iters)
    img = integ.download() / spp_iters
    prim, shad = integ.ray_counts()
    integ.close()
    # a second rank of a 2-rank run renders different samples of the same image
    integ2 = api.Integrator(sc, api.render_settings(flat, "pt", 128, 128, seed=7, rank=1, rank_count=2))
    integ2.render(spp_iters); img2 = integ2.download() / spp_iters
    integ2.close(); sc.close()
    assert 128 * 128 * spp_iters < prim < 128 * 128 * spp_iters * 10 and 0 < shad < prim
    assert np.isfinite(img).all() and img.min() >= 0
    assert not np.array_equal(img, img2) and gamma_rmse(img, img2) < tol_ref  # two independent sample sets (caustic noise in the water scene)
    if refapi.have_ref() and os.path.isdir(refapi.REF_SCENES):
        rs = refapi.RefScene(refapi.scene_path(scene), True)
        ref_img, st = rs.render(0, 128, 128, iterations=64)
        rs.close()
        assert abs(st["pixel_size"] - px) < 1e-4 * px, (st["pixel_size"], px)
        assert gamma_rmse(img, ref_img / 64) < tol_ref
        assert abs(img.mean() - (ref_img / 64).mean()) < 0.02 * img.mean()
    if scene == "cornell_org":
        golden = np.load(os.path.join(GOLDEN, "ref_cornell_org_128.npy"))
        assert gamma_rmse(img, golden) < 0.05


def test_queue_primitives_match_numpy(torch_device=None):
    """compact_hits, sort_by_material, compact_rays (ray_queue.h:161-244) as standalone device calls."""
    import torch
    from imbatracer_b200 import api
    flat = load_scene("cornell_water", "gpu")
    sc = api.Scene(flat)
    L = abi.lib()
    rng = np.random.default_rng(5)
    mats = flat.indices.reshape(-1, 4)[:, 3]
    for n in (0, 1, 31, 2048, 2049, 100003):
        hits = np.zeros(n, abi.HIT)
        hits["tri_id"] = np.where(rng.random(n) < 0.7, rng.integers(0, flat.tri_count, n), -1)
        hits["inst_id"] = 0; hits["tmax"] = rng.random(n); hits["u"] = rng.random(n)
        rays = np.zeros(n, abi.RAY); rays["org"] = rng.random((n, 3)); rays["tmax"] = np.arange(n)
        states = np.zeros(n, abi.PT_STATE); states["pixel_id"] = np.where(rng.random(n) < 0.6, np.arange(n), -1); states["rng"] = rng.integers(0, 2**63, n)
        d_hits = torch.from_numpy(hits.view(np.uint8).reshape(-1).copy()).cuda()
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
        d_states = torch.from_numpy(states.view(np.uint8).reshape(-1).copy()).cuda()
        d_idx = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda"); d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        abi.check(L.imba_queue_sort_by_material(sc.h, d_hits.data_ptr(), n, d_idx.data_ptr(), d_cnt.data_ptr(), None))
        hit_idx = np.nonzero(hits["tri_id"] >= 0)[0]
        want = hit_idx[np.argsort(mats[hits["tri_id"][hit_idx]], kind="stable")]
        assert int(d_cnt.item()) == len(hit_idx)
        assert np.array_equal(d_idx.cpu().numpy()[:len(hit_idx)], want)
        # compact_rays
        o_rays = torch.zeros_like(d_rays); o_states = torch.zeros_like(d_states)
        abi.check(L.imba_queue_compact_rays(d_rays.data_ptr(), d_states.data_ptr(), 40, o_rays.data_ptr(), o_states.data_ptr(), n, d_cnt.data_ptr(), None))
        keep = states["pixel_id"] >= 0
        k = int(d_cnt.item()); assert k == int(keep.sum())
        assert o_rays.cpu().numpy().view(abi.RAY)[:k].tobytes() == rays[keep].tobytes()
        assert o_states.cpu().numpy().view(abi.PT_STATE)[:k].tobytes() == states[keep].tobytes()
        # compact_hits: hits first, both parts in their original order
        o_hits = torch.zeros_like(d_hits)
        abi.check(L.imba_queue_compact_hits(sc.h, d_rays.data_ptr(), d_hits.data_ptr(), d_states.data_ptr(), 40, o_rays.data_ptr(), o_hits.data_ptr(),
                                            o_states.data_ptr(), n, d_cnt.data_ptr(), None))
        order = np.concatenate([hit_idx, np.nonzero(hits["tri_id"] < 0)[0]])
        assert int(d_cnt.item()) == len(hit_idx)
        assert o_hits.cpu().numpy().view(abi.HIT).tobytes() == hits[order].tobytes()
        assert o_rays.cpu().numpy().view(abi.RAY).tobytes() == rays[order].tobytes()
        assert o_states.cpu().numpy().view(abi.PT_STATE).tobytes() == states[order].tobytes()
    sc.close()


def test_headless_driver_prints_reference_report_line_and_matches_golden(tmp_path):
    """python -m imbatracer_b200.render = frontend/main.cpp minus SDL: same flags, same final line, PNG like store_png.h."""
    import re
    import subprocess
    import sys
    from tests.conftest import ROOT
    from PIL import Image
    out = str(tmp_path / "cornell.png")
    golden = np.load(os.path.join(GOLDEN, "ref_cornell_org_128.npy"))
    ref_png = str(tmp_path / "ref.png")
    Image.fromarray((255.0 * np.clip(golden, 0, 1) ** 0.5).astype(np.uint8)).save(ref_png)
    r = subprocess.run([sys.executable, "-m", "imbatracer_b200.render", os.path.join(GOLDEN, "scene_cornell_org_gpu.npz"), out, "-a", "pt", "-w", "128",
                        "-h", "128", "-s", "128", "-q", "--compare", ref_png], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"Done after ([\d.e+-]+) seconds, (\d+) samples @ ([\d.e+-]+) frames per second", r.stdout)  # test/run_bench.py:207
    assert m and int(m.group(2)) == 128
    rmse = float(re.search(r"RMSE ([\d.e+-]+)", r.stdout).group(1))
    assert rmse < 0.05 and os.path.exists(out)


def test_cpp_ray_queue_mirror_against_the_reference_loops(tmp_path):
    """include/imba_b200.hpp RayQueue (ray_queue.h:58-360): thread-safe push from 8 threads, range push, append, and
    compact_hits / sort_by_material / compact_rays delegated to the device (imba_queue_*_host), checked by a C++ program against
    serial restatements of the reference's loops."""
    import subprocess
    from tests.conftest import ROOT
    exe = str(tmp_path / "ray_queue_test")
    lib_dir = os.path.dirname(abi.LIB_PATH)
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "ray_queue_test.cpp"),
                    "-o", exe, "-L", lib_dir, "-limba_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.returncode, r.stdout[-2000:])
