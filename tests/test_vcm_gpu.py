"""GPU parity of the VCM family (LT, BPT, PPM, VCM and the VCM-style PT) against the reference's own
VCMIntegrator<algo> (vcm.cpp) running on the oracle: equal-spp renders must agree in mean radiance (< 1.5 %), in
gamma-0.5 RMSE (the metric of test/run_bench.py:113-129) and in the number of primary / shadow rays per iteration
(the reference's TileScheduler counters, tile_scheduler.h:117-127), and the unbiased algorithms must agree with each
other.  Hash-grid build/query and the light-vertex cache are exercised by PPM / VCM (merging) and BPT / VCM (connections)."""
import os

import numpy as np
import pytest

from tests import refapi
from tests.conftest import load_scene

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refapi.have_ref() or not os.path.isdir(refapi.REF_SCENES), reason="oracle/_ref not built")
REF_ALGO = {"pt": 0, "bpt": 1, "vcm": 2, "ppm": 3, "vcm_pt": 4, "lt": 5}  # UserSettings::Algorithm, frontend/cmd_line.h:41-48


def gamma_rmse(a, b):
    return float(np.sqrt(np.mean((np.clip(a, 0, 1) ** 0.5 - np.clip(b, 0, 1) ** 0.5) ** 2)))


@pytest.fixture(scope="module")
def rendered():
    from imbatracer_b200 import api
    flat = load_scene("cornell_org", "gpu")
    sc = api.Scene(flat)
    out = {}
    n = 48
    for algo in ("pt", "vcm_pt", "lt", "bpt", "ppm", "vcm"):
        it = api.Integrator(sc, api.render_settings(flat, algo, 96, 96, seed=11))
        it.preprocess()
        it.render(n)
        img = it.download() / n
        out[algo] = (img, tuple(c / n for c in it.ray_counts()))
        it.close()
    sc.close()
    return out


def test_algorithms_agree_with_each_other(rendered):
    pt = rendered["pt"][0]
    for algo in ("bpt", "ppm", "vcm"):
        img = rendered[algo][0]
        assert np.isfinite(img).all() and img.min() >= 0
        assert abs(img.mean() - pt.mean()) < 0.02 * pt.mean(), algo
        assert gamma_rmse(img, pt) < 0.04, algo
    # light tracing only sees paths that end at the camera through a connection: darker, never brighter than PT
    lt = rendered["lt"][0]
    assert 0.2 * pt.mean() < lt.mean() < pt.mean()


@needs_ref
@pytest.mark.parametrize("algo,tol", [("vcm_pt", 0.07), ("lt", 0.08), ("bpt", 0.05), ("ppm", 0.05), ("vcm", 0.05)])
def test_matches_reference_integrator(rendered, algo, tol):
    img, (prim, shad) = rendered[algo]
    rs = refapi.RefScene(refapi.scene_path("cornell_org"), True)
    n = 32
    ref, st = rs.render(REF_ALGO[algo], 96, 96, iterations=n, tile=32)  # DefaultTileGen drops slivers < tile/2 (tile_gen.h:51-77)
    rs.close()
    ref = ref / n
    assert abs(img.mean() - ref.mean()) < 0.015 * ref.mean(), (img.mean(), ref.mean())
    assert gamma_rmse(img, ref) < tol, gamma_rmse(img, ref)  # two independent sample sets at 48 / 32 spp
    assert abs(prim - st["primary_rays"] / n) < 0.02 * prim
    assert abs(shad - st["shadow_rays"] / n) <= 0.02 * max(shad, 1)


def test_vcm_needs_preprocess_and_valid_settings():
    from imbatracer_b200 import abi, api
    flat = load_scene("cornell_org", "gpu")
    sc = api.Scene(flat)
    it = api.Integrator(sc, api.render_settings(flat, "vcm", 32, 32))
    with pytest.raises(RuntimeError, match="preprocess"):
        it.render(1)
    it.close()
    with pytest.raises(RuntimeError, match="knn"):
        api.Integrator(sc, api.render_settings(flat, "vcm", 32, 32, num_knn=64))
    sc.close()


@needs_ref
@pytest.mark.parametrize("scene", ["cornell_org", "cornell_specular_front", "cornell_instances", "stilllife_proxy"])
@pytest.mark.parametrize("algo", ["vcm", "bpt", "ppm", "lt", "vcm_pt"])
def test_vcm_stages_match_reference_wavefronts(scene, algo):
    """Stage level: the reference's own process_light_rays / process_camera_rays (vcm.cpp:193-258, 323-439) against the CUDA
    stages on identical injected wavefronts.  Light pass: same survivors in the same order with identical RNG states and path
    lengths, dVC/dVCM/dVM, throughputs and rays within 5e-4; the same camera-connection shadow rays (target pixel exact); the
    same vertex-cache entries (LightVertices::add_vertex_to_cache).  Camera pass with the reference's cache installed on
    both sides: same survivors / states, the same NEE + connection shadow rays (LightVertices::get_connect draws the same
    vertices), and the same framebuffer contributions (emitter hits + HashGrid k-NN merging with the Epanechnikov kernel)."""
    from tests import vcm_harness as vh
    res = vh.run(scene, algo, w=64, h=64, waves=4)
    assert (algo == "vcm_pt") == (len(res["light"]) == 0) and (algo == "lt") == (len(res["camera"]) == 0)
    for r in res["light"]:
        c = r["cont"]
        assert c["ref"] == c["dev"] and c["only_one_side"] == 0, r
        if c["ref"]:
            assert c["same_order"] and c["rng_equal"] == 1.0 and c["len_equal"] == 1.0 and c["close"] >= 0.999, r
        for part in ("connect", "cache"):
            p = r[part]
            assert p["ref"] == p["dev"] and p["only_one_side"] <= 0.002 * max(p["ref"], 1), r
            if p["ref"]:
                assert p["close"] >= 0.998, r
        if algo == "ppm":
            assert r["connect"]["ref"] == 0
        if algo == "lt":
            assert r["cache"]["ref"] == 0
    for r in res["camera"]:
        c = r["cont"]
        assert c["ref"] == c["dev"] and c["only_one_side"] == 0, r
        if c["ref"]:
            assert c["same_order"] and c["rng_equal"] == 1.0 and c["len_equal"] == 1.0 and c["close"] >= 0.999, r
        p = r["shadow"]
        assert p["ref"] == p["dev"] and p["only_one_side"] <= 0.002 * max(p["ref"], 1), r
        if p["ref"]:
            assert p["close"] >= 0.998, r
        assert r["fb_rel_err"] < 2e-3, r


@needs_ref
@pytest.mark.parametrize("algo", ["vcm", "lt", "ppm"])
def test_vcm_stages_with_environment_light(algo):
    """EnvLight::sample_emit for light paths (light.h:494-513, infinite light: finite_light = 0) and the environment term of
    escaped camera paths (vcm.cpp:345-366), through the same stage harness."""
    from tests import vcm_harness as vh
    res = vh.run("cornell_org", algo, w=64, h=64, waves=3, envmap=True)
    for r in res["light"] + res["camera"]:
        c = r["cont"]
        assert c["ref"] == c["dev"] and c["only_one_side"] == 0, r
        if c["ref"]:
            assert c["same_order"] and c["rng_equal"] == 1.0 and c["len_equal"] == 1.0 and c["close"] >= 0.998, r
    for r in res["camera"]:
        assert r["fb_rel_err"] < 2e-3 and r["fb_ref_sum"] > 0, r


def _all_light_kinds_scene():
    """cornell_org plus a point, a spot and a directional light (scene-file grammar of build_scene.cpp:380-450), written next to
    the scratch copy of the Cornell scenes; with the procedural environment map that makes all five light kinds."""
    path = os.path.join(refapi.REF_SCENES, "cornell", "cornell_lights.scene")
    if not os.path.exists(path):
        with open(refapi.scene_path("cornell_org")) as f:
            text = f.read()
        text += "\npoint_light 0.3 1.2 0.1  4 3.5 3\nspot_light -0.4 1.5 0.2  0.2 -1 0.1  35  6 6 5\ndir_light 0.3 -1 -0.2  1.5 1.4 1.2\n"
        with open(path, "w") as f:
            f.write(text)
    return path


@needs_ref
@pytest.mark.parametrize("scene", ["cornell_org", "stilllife_proxy", "all_light_kinds"])
def test_light_path_starts_match_reference_emission_sampling(scene):
    """The first vertex of the light paths: our generator (vcm_light_raygen_kernel: UniformLightTileGen's light choice, seeding,
    light_sample_emit, initial dVCM / dVC / dVM) against the reference's Light::sample_emit (light.h) driven with the same
    (light, RNG state) pairs -- triangle lights, and in the third scene also point, spot, directional and environment lights."""
    from imbatracer_b200 import abi, api
    if scene == "all_light_kinds":
        rs = refapi.RefScene(_all_light_kinds_scene(), True)
        rs.add_test_envmap(5, 0.7)
    else:
        rs = refapi.RefScene(refapi.scene_path(scene), True)
    flat = rs.flat(scene)
    kinds = set(int(k) for k in flat.lights["kind"])
    assert kinds == ({0, 1, 2, 3, 4} if scene == "all_light_kinds" else {0})
    sc = api.Scene(flat)
    n = 20000
    light_paths = 30000
    it = api.Integrator(sc, api.render_settings(flat, "vcm", 64, 64, light_path_count=light_paths, seed=5))
    it.preprocess()
    base_radius, iteration, seed_base = 0.02, 3, 0x51F15EED
    rays, states = it.vcm_raygen_batch(True, iteration, base_radius, seed_base, n)
    it.close(); sc.close()
    rng0 = np.zeros(n, np.uint64)
    abi.check(abi.lib().imba_seed_states(seed_base, 0, n, 1, rng0.ctypes.data))
    f32 = np.float32
    pm_radius = max(f32(base_radius) / f32(np.power(f32(iteration), f32(0.5 * (1.0 - 0.75)))), f32(1e-7))
    mis_eta_vc = float(f32(1.0) / (f32(np.pi) * pm_radius * pm_radius * f32(light_paths)))
    ref_rays, ref_states = rs.light_emit(states["light_id"], rng0, mis_eta_vc)
    rs.close()
    # every light gets its share of the paths (UniformLightTileGen, tile_gen.h:139-158)
    assert set(states["light_id"].tolist()) == set(range(len(flat.lights)))
    assert np.array_equal(states["rng"], ref_states["rng"])                       # same number of random draws, same order
    assert np.array_equal(states["len_finite"], ref_states["len_finite"])         # path length 1 + finite-light flag

    def close(a, b, tol=2e-5):
        a, b = a.astype(np.float64), b.astype(np.float64)
        return np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    assert close(rays["org"], ref_rays["org"]).all() and close(rays["dir"], ref_rays["dir"]).all()
    assert np.array_equal(rays["tmin"], ref_rays["tmin"]) and np.array_equal(rays["tmax"], ref_rays["tmax"])
    for k in ("throughput", "dVCM", "dVC"):
        assert close(states[k], ref_states[k], 1e-4).all(), k
    assert close(states["dVM"], ref_states["dVM"], 1e-3).all()                    # carries mis_eta_vc, recomputed here in numpy


@needs_ref
def test_camera_path_starts_match_reference():
    """The first vertex of the camera paths of the VCM family (vcm_camera_raygen_kernel) against the sample function of the
    reference's trace_camera_paths (vcm.cpp:109-129): jittered raster position, camera ray, throughput 1, dVC = dVM = 0 and
    dVCM = light_path_count / pdf_cam_w."""
    from imbatracer_b200 import abi, api
    w, h, light_paths, seed_base = 80, 60, 12345, 0x0C0FFEE1
    rs = refapi.RefScene(refapi.scene_path("cornell_org"), True)
    flat = rs.flat("cornell_org")
    sc = api.Scene(flat)
    it = api.Integrator(sc, api.render_settings(flat, "vcm", w, h, light_path_count=light_paths, seed=5))
    it.preprocess()
    rays, states = it.vcm_raygen_batch(False, 2, 0.02, seed_base, w * h)
    it.close(); sc.close()
    rng0 = np.zeros(w * h, np.uint64)
    abi.check(abi.lib().imba_seed_states(seed_base, 0, w * h, 1, rng0.ctypes.data))
    ref_rays, ref_states = rs.camera_start(w, h, light_paths, states["pixel_id"], rng0)
    rs.close()
    assert np.array_equal(states["pixel_id"], np.arange(w * h))
    assert np.array_equal(states["rng"], ref_states["rng"]) and np.array_equal(states["len_finite"], ref_states["len_finite"])
    assert np.array_equal(rays["org"], ref_rays["org"]) and np.abs(rays["dir"] - ref_rays["dir"]).max() < 5e-6
    assert np.array_equal(states["throughput"], ref_states["throughput"])
    assert (states["dVC"] == 0).all() and (states["dVM"] == 0).all()
    assert np.allclose(states["dVCM"], ref_states["dVCM"], rtol=2e-5, atol=0)


@pytest.mark.gpu
def test_reset_restarts_the_progressive_sequence():
    """Integrator::reset (integrator.h:31; VCMIntegrator: cur_iteration_ = 0, vcm.h:69-72): after clear() + reset() the same
    iterations are rendered again -- same seeds, same PPM radius schedule -- so ray counts and image repeat."""
    from imbatracer_b200 import api
    flat = load_scene("cornell_org", "gpu")
    sc = api.Scene(flat)
    it = api.Integrator(sc, api.render_settings(flat, "ppm", 96, 96, seed=8))
    it.preprocess()
    it.render(3)
    a, ca = it.download(), it.ray_counts()
    it.render(2)                                   # iterations 3 and 4 have smaller radii: a different image per iteration
    b = it.download() - a
    it.clear(); it.reset()
    it.render(3)
    c, cc = it.download(), it.ray_counts()
    it.close(); sc.close()
    assert cc == ca
    assert np.abs(c - a).max() <= 1e-3 * max(1.0, a.max()) and abs(c.sum() - a.sum()) <= 2e-4 * a.sum()
    assert abs(b.sum() / 2 - a.sum() / 3) > 1e-6 * a.sum()     # not trivially equal images


@needs_ref
@pytest.mark.parametrize("scene", ["cornell_org", "cornell_indirect", "cornell_water", "cornell_instances", "cornell_specular_front", "stilllife_proxy"])
def test_light_vertex_cache_size_matches_reference_probe(scene):
    """LightVertices::compute_cache_size (light_vertices.cpp:15-129): both sides trace 10 000 probe light paths and allocate
    margin * ceil(average number of stored vertices per path) * light_path_count slots, margin = 1.1, or 10 for fewer than 1000
    paths (:125-126).  The average is a Monte-Carlo estimate, the ceil() makes the result an integer step function of it: the slot
    counts must be EQUAL (the six scenes' averages lie between 2.x and 7.x, none close to an integer)."""
    from imbatracer_b200 import abi, api
    flat = load_scene(scene, "gpu")
    rs = refapi.RefScene(refapi.scene_path(scene), True)
    sc = api.Scene(flat)
    L = abi.lib()
    try:
        for paths in (500, 4608, 131072):
            want = rs.vcm_cache_size(paths)
            for algo in ("vcm", "bpt", "ppm"):
                it = api.Integrator(sc, api.render_settings(flat, algo, 96, 96, light_path_count=paths, seed=21))
                it.preprocess()
                got = L.imba_integrator_vertex_cache_size(it.h)
                it.close()
                assert got == want, (scene, algo, paths, got, want)
        for algo in ("lt", "pt"):   # no cache: LT splats directly, PT has no light pass (vcm.h:78-80)
            it = api.Integrator(sc, api.render_settings(flat, algo, 96, 96, light_path_count=4608))
            it.preprocess()
            assert L.imba_integrator_vertex_cache_size(it.h) == 0
            it.close()
    finally:
        sc.close(); rs.close()


@needs_ref
@pytest.mark.parametrize("algo", ["vcm", "bpt", "lt"])
def test_vcm_stages_with_phong_materials(algo):
    """Phong (brdfs.h:48-92) in the bidirectional family: light-vertex storage on a glossy-reflection BSDF, connections evaluating it in
    both directions (forward / reverse pdfs of the MIS partial weights), adjoint sampling on light paths -- same stage harness."""
    from tests import vcm_harness as vh
    res = vh.run("cornell_org", algo, w=64, h=64, waves=3, phong=True)
    for r in res["light"] + res["camera"]:
        c = r["cont"]
        assert c["ref"] == c["dev"] and c["only_one_side"] == 0, r
        if c["ref"]:
            assert c["same_order"] and c["rng_equal"] == 1.0 and c["len_equal"] == 1.0 and c["close"] >= 0.998, r
    for r in res["light"]:
        for part in ("connect", "cache"):
            p = r[part]
            assert p["ref"] == p["dev"] and p["only_one_side"] <= 0.002 * max(p["ref"], 1), r
            if p["ref"]:
                assert p["close"] >= 0.998, r
    for r in res["camera"]:
        assert r["fb_rel_err"] < 2e-3, r
