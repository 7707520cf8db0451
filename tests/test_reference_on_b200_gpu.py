"""The UNMODIFIED REFERENCE running on the product library.

oracle/_ref/libimba_ref_b200.so (oracle/Makefile: ref_b200) is the reference's own host code -- loader, SBVH builder, adapters,
Scene upload, RayQueue, TileScheduler, PathTracer, VCMIntegrator -- compiled from /root/reference/src exactly like
libimba_ref.so, but linked against imbatracer_b200/libimba_b200.so INSTEAD OF the traversal restatement (the link line of
src/imbatracer/CMakeLists.txt:76-87), with the AnyDSL runtime shim backed by the CUDA runtime so that `--gpu` scenes really
live in device memory.  So RayQueue::traverse_{cpu,gpu} / traverse_occluded_{cpu,gpu} (ray_queue.h:247-343) call the product's
four entry points from TileScheduler's worker threads (tile_scheduler.h:69-78,120-132), with anydsl::Array device pointers,
padded counts, per-queue H2D / D2H copies -- everything a ctypes call cannot exercise.

With std::random_device made reproducible (oracle/shim/det_random.h) and ONE scheduler thread the two builds trace the same
rays: ray counters must be IDENTICAL and the images equal up to the order of float additions.  With four scheduler threads
the tile -> seed assignment races, so that case is compared statistically.

Second part: the product's plug-ins AT the reference's interfaces (include/imba_b200_ref.hpp): imba::B200Integrator, a subclass
of imba::Integrator, and imba::B200WavefrontScheduler, a subclass of imba::RayScheduler, constructed and driven the way
frontend/main.cpp:117-163 and render_window.cpp:93-105 drive the reference's own classes."""
import os

import numpy as np
import pytest

from tests import refapi

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refapi.have_ref() and refapi.have_ref_b200() and os.path.isdir(refapi.REF_SCENES)), reason="oracle/_ref not built")]
ALGO = {"pt": 0, "bpt": 1, "vcm": 2, "ppm": 3, "vcm_pt": 4, "lt": 5}   # UserSettings::Algorithm, frontend/cmd_line.h:41-48


def pooled_rmse(a, b, k=8):
    """Gamma-0.5 RMSE of the k x k box-filtered images: two INDEPENDENT 24-32 spp renders of the reference itself differ by 0.21 per
    pixel on cornell_instances (BPT, 42 lights, glass) but by 0.028-0.037 after 8 x 8 pooling (cornell_org: 0.004-0.005), so the
    pooled figure separates Monte-Carlo noise from a wrong image.  Tolerances below = 2-3 x those calibration values."""
    h, w, _ = a.shape
    pa = a.reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3)); pb = b.reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))
    return float(np.sqrt(np.mean((np.clip(pa, 0, 1) ** 0.5 - np.clip(pb, 0, 1) ** 0.5) ** 2)))


TOL = {"cornell_org": (0.015, 0.02), "cornell_instances": (0.08, 0.04), "cornell_water": (0.04, 0.03)}   # pooled RMSE, relative mean


def render_both(scene, algo, gpu_layout, threads, iterations, seed, size=96, tile=32):
    out = []
    for lib in (refapi.ref(), refapi.ref_b200()):
        rs = refapi.RefScene(refapi.scene_path(scene), gpu_layout, lib=lib)
        lib.ref_set_seed(seed)
        img, st = rs.render(ALGO[algo], size, size, iterations=iterations, threads=threads, tile=tile)
        lib.ref_set_seed(0)
        rs.close()
        out.append((img / iterations, st))
    return out


@pytest.mark.parametrize("scene,algo,gpu_layout", [("cornell_org", "pt", False), ("cornell_org", "pt", True), ("cornell_instances", "pt", True),
                                                  ("cornell_instances", "bpt", False), ("cornell_instances", "bpt", True)])
def test_reference_traces_the_same_rays_on_the_library(scene, algo, gpu_layout):
    """(VCM / PPM are left to the statistical test: their merging and connection partners come from the light-vertex cache, whose
    order is the arrival order of atomic appends from the parallel shading loop and differs from run to run in ONE build already.)"""
    (ref_img, ref_st), (img, st) = render_both(scene, algo, gpu_layout, threads=1, iterations=3, seed=0xC0FFEE)
    assert st["primary_rays"] == ref_st["primary_rays"] and st["shadow_rays"] == ref_st["shadow_rays"], (st, ref_st)
    assert st["primary_rays"] > 3 * 96 * 96 and st["shadow_rays"] > 0
    assert abs(st["pixel_size"] - ref_st["pixel_size"]) <= 1e-6 * ref_st["pixel_size"]
    # same paths, same contributions; only the order of the float additions into a pixel may differ (parallel shading loops)
    assert np.abs(img - ref_img).sum() <= 1e-4 * np.abs(ref_img).sum(), float(np.abs(img - ref_img).sum() / np.abs(ref_img).sum())


@pytest.mark.parametrize("scene,algo,gpu_layout", [("cornell_org", "pt", True), ("cornell_instances", "bpt", False), ("cornell_instances", "bpt", True),
                                                  ("cornell_water", "vcm", True)])
def test_reference_with_four_scheduler_threads_on_the_library(scene, algo, gpu_layout):
    """thread_count = 4: TileScheduler's workers call the entry points concurrently on disjoint queues with the shared scene."""
    (ref_img, ref_st), (img, st) = render_both(scene, algo, gpu_layout, threads=4, iterations=24, seed=0)
    assert np.isfinite(img).all() and img.min() >= 0
    tol, mean_tol = TOL[scene]
    assert abs(img.mean() - ref_img.mean()) < mean_tol * ref_img.mean(), (img.mean(), ref_img.mean())
    assert pooled_rmse(img, ref_img) < tol, pooled_rmse(img, ref_img)
    assert abs(st["primary_rays"] - ref_st["primary_rays"]) < 0.02 * ref_st["primary_rays"]
    assert abs(st["shadow_rays"] - ref_st["shadow_rays"]) < 0.02 * ref_st["shadow_rays"]


def test_a_second_scene_at_the_same_addresses_is_not_traversed_with_the_stale_tree():
    """The entry points receive raw arrays; the converted structure is cached per (addresses, root, extents, content fingerprint).
    Loading, rendering and freeing scene after scene in one process makes cudaMalloc hand the same addresses out again."""
    L = refapi.ref_b200()
    means = {}
    for rnd in range(2):
        for scene in ("cornell_org", "cornell_indirect", "cornell_specular_front"):
            rs = refapi.RefScene(refapi.scene_path(scene), True, lib=L)
            L.ref_set_seed(5)
            img, st = rs.render(ALGO["pt"], 64, 64, iterations=2, threads=2, tile=32)
            L.ref_set_seed(0)
            rs.close()
            means.setdefault(scene, []).append((float(img.mean()), st["primary_rays"]))
    for scene, v in means.items():
        assert v[0][1] == v[1][1] and abs(v[0][0] - v[1][0]) <= 1e-4 * v[0][0], (scene, v)   # the same render both times
    assert len({round(v[0][0], 4) for v in means.values()}) == 3                                   # and three different scenes
    rs = refapi.RefScene(refapi.scene_path("cornell_org"), True, lib=refapi.ref())
    refapi.ref().ref_set_seed(5)
    img, st = rs.render(ALGO["pt"], 64, 64, iterations=2, threads=2, tile=32)
    refapi.ref().ref_set_seed(0); rs.close()
    assert abs(float(img.mean()) - means["cornell_org"][1][0]) < 0.05 * float(img.mean())


@pytest.mark.parametrize("scene,algo", [("cornell_org", "pt"), ("cornell_instances", "bpt"), ("cornell_org", "vcm")])
def test_b200_integrator_is_an_imba_integrator(scene, algo):
    """imba::B200Integrator : public imba::Integrator -- render(AtomicImage&) adds one un-normalised iteration (render_window.cpp:105),
    preprocess() gives the reference's pixel_size() through the linked symbols and the device's own probe agrees with it."""
    n = 32
    rs = refapi.RefScene(refapi.scene_path(scene), False, lib=refapi.ref_b200())
    img, st = rs.render_plugin(1, ALGO[algo], 96, 96, iterations=n, seed=9)
    rs.close()
    rs = refapi.RefScene(refapi.scene_path(scene), False)
    ref_img, ref_st = rs.render(ALGO[algo], 96, 96, iterations=n, tile=32)
    rs.close()
    img /= n; ref_img /= n
    assert abs(st["pixel_size"] - ref_st["pixel_size"]) <= 1e-5 * ref_st["pixel_size"]
    assert abs(st["device_pixel_size"] - ref_st["pixel_size"]) <= 1e-4 * ref_st["pixel_size"]
    tol, mean_tol = TOL[scene]
    assert abs(img.mean() - ref_img.mean()) < mean_tol * ref_img.mean(), (img.mean(), ref_img.mean())
    assert pooled_rmse(img, ref_img) < tol, pooled_rmse(img, ref_img)
    assert abs(st["primary_rays"] - ref_st["primary_rays"]) < 0.02 * ref_st["primary_rays"], (st, ref_st)
    assert abs(st["shadow_rays"] - ref_st["shadow_rays"]) <= 0.02 * max(ref_st["shadow_rays"], 1), (st, ref_st)


@pytest.mark.parametrize("scene,algo,gpu_layout", [("cornell_org", "pt", True), ("cornell_org", "pt", False), ("cornell_instances", "bpt", True)])
def test_b200_wavefront_scheduler_is_an_imba_ray_scheduler(scene, algo, gpu_layout):
    """imba::B200WavefrontScheduler : public imba::RayScheduler -- the reference's PathTracer / VCMIntegrator run unchanged on it: one
    wavefront over the whole frame instead of tiles, traversal on the device, the reference's shading callbacks."""
    n = 24
    rs = refapi.RefScene(refapi.scene_path(scene), gpu_layout, lib=refapi.ref_b200())
    img, st = rs.render_plugin(0, ALGO[algo], 96, 96, iterations=n)
    rs.close()
    rs = refapi.RefScene(refapi.scene_path(scene), gpu_layout)
    ref_img, ref_st = rs.render(ALGO[algo], 96, 96, iterations=n, tile=32)
    rs.close()
    img /= n; ref_img /= n
    assert abs(st["pixel_size"] - ref_st["pixel_size"]) <= 1e-5 * ref_st["pixel_size"]
    tol, mean_tol = TOL[scene]
    assert abs(img.mean() - ref_img.mean()) < mean_tol * ref_img.mean(), (img.mean(), ref_img.mean())
    assert pooled_rmse(img, ref_img) < tol, pooled_rmse(img, ref_img)
    if algo == "pt":   # (the VCM family counts its light pass on a scheduler of its own)
        assert abs(st["primary_rays"] - ref_st["primary_rays"]) < 0.02 * ref_st["primary_rays"]
        assert abs(st["shadow_rays"] - ref_st["shadow_rays"]) < 0.02 * ref_st["shadow_rays"]
