#!/usr/bin/env python
"""bench.py -- traversal micro-bench of BASELINE.json configs[1]: 4 Mi coherent primary rays and 4 Mi incoherent
diffuse-bounce rays, closest-hit and occlusion, per step (16 Mi rays), on the BVH of "stilllife-proxy": the Still Life
geometry is not shipped with the reference, so the scene is a procedural stand-in with the camera of
still_life.scene and the 23 materials of still_life.mtl (imbatracer_b200/procedural.py, ~1.5 M triangles, structure built
by the product's host builder in the reference's BVH2 layout so that the CPU arm traces the very same tree).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` times the device-resident path (imba_trace on HBM-resident queues), `e2e`
times the reference-facing host-buffer call (imba_trace_host = RayQueue::traverse_gpu: H2D rays, trace, D2H hits).
`--impl reference` times the CPU restatement of the reference's traversal entry points (oracle/, all host threads)
on a bounded sample of the same batches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_BATCH = 1 << 22
SCENE = "stilllife-proxy"
LAYOUT = "gpu"  # the reference's BVH2 ("--gpu") layout converts to the better 8-wide tree
PROXY_DETAIL = float(os.environ.get("IMBA_BENCH_DETAIL", "1.0"))
METRIC = "Mrays/s (closest-hit + occlusion, coherent + incoherent)"


def load_flat():
    from imbatracer_b200 import procedural
    return procedural.still_life_proxy(PROXY_DETAIL, name=SCENE, accel="host")


def load_cornell(name="cornell_water"):
    from imbatracer_b200.scene import FlatScene
    return FlatScene.load(os.path.join(ROOT, "tests", "golden", f"scene_{name}_{LAYOUT}.npz"), name)


def camera_basis(flat, fov=60.0):
    pos, d, up = flat.cam[0:3].astype(np.float64), flat.cam[3:6].astype(np.float64), flat.cam[6:9].astype(np.float64)
    d /= np.linalg.norm(d); up /= np.linalg.norm(up)
    right = np.cross(d, up); right /= np.linalg.norm(right)
    up = np.cross(right, d)
    return pos, d, right, up, np.tan(np.radians(fov) / 2)


def make_batches(flat, n, seed, scene=None, xp=None):
    """Synthetic, seeded input queues (not timed): camera rays through a sqrt(n) x sqrt(n) jittered pixel grid; their
    cosine-weighted diffuse bounces (origin = hit point, tmin = 1e-3 t, pt.cpp:76-79,131); shadow rays from the primary
    and from the bounce hit points towards points on the emitters (tmax = 1 - 1e-3, pt.cpp:37-41).
    With `scene` the hits come from the CUDA path; without (reference arm) from the oracle on the host."""
    from imbatracer_b200 import abi
    rng = np.random.default_rng(seed)
    side = int(np.sqrt(n)); assert side * side == n
    pos, d, right, up, tan = camera_basis(flat)
    px = (np.arange(n) % side + rng.random(n)) / side * 2 - 1
    py = 1 - (np.arange(n) // side + rng.random(n)) / side * 2
    dirs = d[None, :] + (px * tan)[:, None] * right[None, :] + (py * tan)[:, None] * up[None, :]
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    prim = np.zeros(n, abi.RAY); prim["org"] = pos.astype(np.float32); prim["dir"] = dirs.astype(np.float32)
    prim["tmax"] = np.finfo(np.float32).max

    def trace(r):
        if scene is not None:
            return scene.trace_host(r)
        from tests import refapi
        return refapi.trace_oracle(flat, r)

    gn = flat.geom_normals.reshape(-1, 4)[:, :3]
    lights = flat.lights

    def bounce_and_shadow(rays, hits):
        t = hits["tmax"].astype(np.float64)[:, None]
        hit = hits["tri_id"] >= 0
        p = rays["org"].astype(np.float64) + t * rays["dir"].astype(np.float64)
        nrm = gn[np.clip(hits["tri_id"], 0, len(gn) - 1)].astype(np.float64)
        nrm *= np.where((nrm * rays["dir"]).sum(1, keepdims=True) > 0, -1.0, 1.0)
        # cosine-weighted hemisphere about the facing normal
        u1, u2 = rng.random(len(rays)), rng.random(len(rays))
        r_, phi = np.sqrt(u1), 2 * np.pi * u2
        a = np.where(np.abs(nrm[:, :1]) > 0.9, [[0.0, 1.0, 0.0]], [[1.0, 0.0, 0.0]])
        tx = np.cross(nrm, a); tx /= np.linalg.norm(tx, axis=1, keepdims=True); ty = np.cross(nrm, tx)
        bd = tx * (r_ * np.cos(phi))[:, None] + ty * (r_ * np.sin(phi))[:, None] + nrm * np.sqrt(1 - u1)[:, None]
        b = np.zeros(len(rays), abi.RAY)
        b["org"] = np.where(hit[:, None], p, rays["org"]).astype(np.float32)
        b["dir"] = np.where(hit[:, None], bd, rays["dir"]).astype(np.float32)
        b["tmin"] = np.where(hit, 1e-3 * t[:, 0], 0).astype(np.float32); b["tmax"] = np.finfo(np.float32).max
        # shadow ray to a uniformly sampled point of a uniformly picked triangle light; dir un-normalised, tmax < 1
        li = rng.integers(0, len(lights), len(rays))
        s1, s2 = rng.random(len(rays)), rng.random(len(rays))
        su = 1 - np.sqrt(s1); sv = s2 * np.sqrt(s1)
        lp = (lights["p0"][li, :3] * (1 - su - sv)[:, None] + lights["p1"][li, :3] * su[:, None] + lights["p2"][li, :3] * sv[:, None]).astype(np.float64)
        s = np.zeros(len(rays), abi.RAY)
        s["org"] = b["org"]; s["dir"] = (lp - b["org"]).astype(np.float32)
        s["tmin"] = 1e-3; s["tmax"] = 1 - 1e-3
        return b, s

    h0 = trace(prim)
    bounce, shadow0 = bounce_and_shadow(prim, h0)
    h1 = trace(bounce)
    _, shadow1 = bounce_and_shadow(bounce, h1)
    return {"coherent_closest": prim, "incoherent_closest": bounce, "coherent_occluded": shadow0, "incoherent_occluded": shadow1}


BATCH_ORDER = [("coherent_closest", 0), ("incoherent_closest", 0), ("coherent_occluded", 1), ("incoherent_occluded", 1)]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     stdout=subprocess.PIPE, text=True, timeout=5).stdout.strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm, reasons, mx = [], set(), None
        for s in self.samples:
            if len(s) < 6:
                continue
            try:
                sm.append(float(s[0])); mx = float(s[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """SHA-256 (first 16 hex digits) of the sources the traversal kernel and its data layout are compiled from: what a committed
    ncu capture must have been taken on to describe the kernel this run times."""
    import hashlib
    h = hashlib.sha256()
    for f in ("traverse.cu", "traverse_core.cuh", "bvh8.h", "bvh8_build.cpp"):
        h.update(open(os.path.join(ROOT, "imbatracer_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the incoherent closest-hit launch, from the last `ncu --set full` capture of
    this workload (tools/ncu_summary.py writes profiles/ncu_traffic.json and stamps it with kernel_source_hash()).  A capture taken
    on other kernel sources or another workload is REFUSED (None): a stale constant would go unnoticed otherwise."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no capture committed"
    t = json.load(open(p))
    if t.get("workload") != SCENE:
        return None, f"capture is of workload {t.get('workload')!r}"
    if t.get("kernel_source_hash") != kernel_source_hash():
        return None, f"capture was taken on kernel sources {t.get('kernel_source_hash')}, this tree is {kernel_source_hash()}: re-run tools/gpu_round.sh full"
    return t.get("dram_bytes_per_launch"), f"{t.get('source')} (kernel sources {t.get('kernel_source_hash')})"


def cpu_arm(flat, batches, sample, threads):
    """The oracle (CPU restatement of the reference's traversal entry points) on `sample` rays of every batch."""
    from tests import refapi
    total, secs = 0, 0.0
    for name, any_hit in BATCH_ORDER:
        r = np.ascontiguousarray(batches[name][:sample])
        t0 = time.perf_counter()
        refapi.trace_oracle(flat, r, any_hit=bool(any_hit), threads=threads)
        secs += time.perf_counter() - t0
        total += len(r)
    return total, secs


CFG0 = {"width": 512, "height": 512, "spp": 16, "max_path_len": 22, "tile": 256}   # BASELINE.json configs[0]; path length as test/run_bench.py:75


def reference_cfg0(cores):
    """configs[0] on the host: the reference's OWN PathTracer + TileScheduler + ray queues (unmodified sources compiled into
    oracle/_ref/libimba_ref.so, TBB shimmed onto std::thread) with the restated CPU traversal behind its four entry points, on
    the detail-0.02 stilllife-proxy scene loaded by the reference's own OBJ loader and SBVH builder (MBVH4, its --cpu path)."""
    from tests import refapi
    if not refapi.have_ref():
        return {"unavailable": "oracle/_ref/libimba_ref.so not built"}
    rs = refapi.RefScene(refapi.scene_path("stilllife_proxy"), False)
    rs.render(0, CFG0["width"], CFG0["height"], spp=1, iterations=1, threads=cores, tile=CFG0["tile"], max_path_len=CFG0["max_path_len"])
    img, st = rs.render(0, CFG0["width"], CFG0["height"], spp=1, iterations=CFG0["spp"], threads=cores, tile=CFG0["tile"], max_path_len=CFG0["max_path_len"])
    rs.close()
    n = CFG0["width"] * CFG0["height"] * CFG0["spp"]
    return {"workload": "PT 512x512, 16 spp, max path length 22, tile 256, stilllife-proxy at detail 0.02 (31 k triangles) through the reference's loader",
            "samples_per_s": n / st["seconds"], "seconds": st["seconds"], "mrays_s": (st["primary_rays"] + st["shadow_rays"]) / st["seconds"] / 1e6,
            "threads": cores, "mean_radiance": float(img.mean()) / CFG0["spp"], "kind": "reference host pipeline + restated traversal"}


def reference_on_library_cfg0(cores):
    """configs[0] once more, this time THE UNMODIFIED REFERENCE LINKED AGAINST libimba_b200.so (oracle/_ref/libimba_ref_b200.so, oracle/Makefile:
    ref_b200): its own loader, SBVH builder, TileScheduler, RayQueue and PathTracer, with the four traversal symbols resolved from the product
    and the scene in device memory (`--gpu`): every 65 536-ray queue is copied up, traversed and copied down by the reference's own code
    (ray_queue.h:269-343).  What a maintainer gets from the zero-source-change drop-in; the shading stays on the host's cores."""
    from tests import refapi
    if not refapi.have_ref_b200():
        return {"unavailable": "oracle/_ref/libimba_ref_b200.so not built"}
    rs = refapi.RefScene(refapi.scene_path("stilllife_proxy"), True, lib=refapi.ref_b200())
    rs.render(0, CFG0["width"], CFG0["height"], spp=1, iterations=1, threads=cores, tile=CFG0["tile"], max_path_len=CFG0["max_path_len"])
    img, st = rs.render(0, CFG0["width"], CFG0["height"], spp=1, iterations=CFG0["spp"], threads=cores, tile=CFG0["tile"], max_path_len=CFG0["max_path_len"])
    rs.close()
    n = CFG0["width"] * CFG0["height"] * CFG0["spp"]
    return {"workload": "PT 512x512, 16 spp, max path length 22, tile 256 (65 536-ray queues), stilllife-proxy at detail 0.02, the reference's --gpu path",
            "samples_per_s": n / st["seconds"], "seconds": st["seconds"], "mrays_s": (st["primary_rays"] + st["shadow_rays"]) / st["seconds"] / 1e6,
            "threads": cores, "mean_radiance": float(img.mean()) / CFG0["spp"],
            "kind": "unmodified reference host pipeline, traversal entry points bound to libimba_b200.so (queues cross PCIe per call, shading on the host)"}


def ours_cfg0():
    """The same configuration on the device integrator, on the same scene (the fixture the reference's loader exported)."""
    import torch
    from imbatracer_b200 import api
    from imbatracer_b200.scene import FlatScene
    flat = FlatScene.load(os.path.join(ROOT, "tests", "golden", "scene_stilllife_proxy_gpu.npz"), "stilllife_proxy")
    sc = api.Scene(flat)
    it = api.Integrator(sc, api.render_settings(flat, "pt", CFG0["width"], CFG0["height"], spp=1, max_path_len=CFG0["max_path_len"], seed=7,
                                                 iterations_per_pass=CFG0["spp"]))   # the 16 iterations as one wavefront pass of 4 Mi paths
    it.preprocess(); it.render(CFG0["spp"]); it.sync(); it.clear()
    t0 = time.perf_counter(); it.render(CFG0["spp"]); it.sync(); dt = time.perf_counter() - t0
    img = it.download()
    prim, shad = it.ray_counts()
    it.close(); sc.close()
    n = CFG0["width"] * CFG0["height"] * CFG0["spp"]
    return {"workload": "PT 512x512, 16 spp, max path length 22, stilllife-proxy at detail 0.02 (31 k triangles), same scene arrays as the reference arm",
            "samples_per_s": n / dt, "seconds": dt, "mean_radiance": float(img.mean()) / CFG0["spp"]}


def reference_built_flat():
    """The bench scene with the acceleration structure the REFERENCE builds for it: the procedural mesh is written as OBJ / MTL /
    TGA / .scene (tests/obj_export.py), read back by the reference's own loader and built by its SplitBvhBuilder + GPU adapter
    (unmodified sources in oracle/_ref/libimba_ref.so; ~25 s for 1.5 M triangles, cached as .bvh by the loader itself,
    render/scene.cpp:67).  Returns (flat scene in the reference's BVH2 layout, description) or (None, reason)."""
    from tests import refapi
    if not refapi.have_ref():
        return None, "oracle/_ref/libimba_ref.so not built"
    from imbatracer_b200 import procedural
    from tests import obj_export
    d = os.path.join(refapi.REF_SCENES, f"stilllife_proxy_d{PROXY_DETAIL:g}")
    scene_file = os.path.join(d, "stilllife_proxy.scene")
    if not os.path.exists(scene_file):
        os.makedirs(refapi.REF_SCENES, exist_ok=True)
        obj_export.write_scene(procedural.still_life_proxy(PROXY_DETAIL, accel=None), d)
    t0 = time.perf_counter()
    rs = refapi.RefScene(scene_file, True)
    flat = rs.flat(SCENE)
    rs.close()
    return flat, f"reference loader + SplitBvhBuilder + gpu_adapter (BVH2), {time.perf_counter() - t0:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build()
    flat, tree = reference_built_flat()
    if flat is None:   # no reference build on this box: the product's host builder, same layout
        flat, tree = load_flat(), "product host builder (binned SAH, BVH2 layout): " + tree
    cores = os.cpu_count()
    n = RAYS_PER_BATCH  # the full workload: about 4 s of host time per step
    batches = make_batches(flat, n, 0x1234ABCD)
    for _ in range(args.warmup):
        cpu_arm(flat, batches, 1 << 16, cores)
    total, secs = 0, 0.0
    for _ in range(args.steps):
        a, b = cpu_arm(flat, batches, n, cores); total += a; secs += b
    v = total / secs / 1e6
    cfg0 = reference_cfg0(cores)
    line = {"impl": "reference", "cfg0": cfg0, "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(flat, n),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"4 batches x {n} rays per step; traversal = oracle/traversal_oracle.cpp, a SCALAR std::thread restatement of the four entry "
                                       f"points (the AnyDSL kernels are not in the reference tree); acceleration structure = {tree}"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(flat, n):
    return {"workload": f"traversal micro-bench on the {SCENE} BVH (procedural stand-in for Still Life, whose geometry is absent from the reference "
                        f"tree: camera of still_life.scene, the 23 materials of still_life.mtl, {flat.tri_count} triangles, reference BVH2 layout): per step {n} coherent primary + {n} diffuse-bounce closest-hit rays and "
                        f"{n} + {n} shadow rays", "rays_per_step": 4 * n, "l2_policy": "inputs larger than L2 (each ray queue is 128 MiB, each hit queue 64 MiB)",
            "parallelism": "independent ray queues per GPU, no collective (value, e2e); PT: progressive iterations sharded over the ranks, one NCCL sum of the framebuffer per step (config.pt_*)"}


def run_pt(scene, flat, rank, world, barrier, dist, iters, width=1920, height=1080, spp=1, per_pass=1, overlap=None):
    """configs[2]: progressive PT, iterations sharded over the ranks (rank r renders iterations r, r + R, ...), one NCCL sum of the
    W x H x 3 float framebuffer per step of `per_pass` iterations.  `iters` = iterations PER RANK inside the timed region.
    overlap=None: the sum of step k runs on a second stream beside the rendering of step k + 1 (the default of ProgressiveRenderer
    with more than one rank); overlap=False: render, sum, render, sum ... on one stream."""
    import torch
    from imbatracer_b200 import api
    integ = api.Integrator(scene, api.render_settings(flat, "pt", width, height, spp=spp, max_path_len=10, seed=0x1234ABCD, rank=rank, rank_count=world, iterations_per_pass=per_pass))
    from imbatracer_b200.distributed import ProgressiveRenderer
    integ.preprocess()
    loop = ProgressiveRenderer(integ, width, height, spp, iterations_per_step=per_pass, overlap=overlap)   # integ.render(per_pass) + NCCL sum of the framebuffer per step
    stream = loop.stream
    warm = 2 * per_pass
    loop.step(warm)
    loop.finish(); integ.sync(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    loop.step(iters)
    loop.finish()                      # the framebuffer sums run on a second stream; the timed region ends after the last one
    e1.record(stream)
    integ.sync(); barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    total = loop.total
    prim, shad = integ.ray_counts()
    mean = float(total.mean()) / ((iters + warm) * world * spp)
    integ.close()
    samples = width * height * spp * iters * world
    return {"workload": f"PT {width}x{height}, {spp} spp per iteration per GPU (the reference's --spp), max path length 10, {SCENE}", "iterations_per_rank": iters,
            "total_spp": iters * world * spp, "seconds": ms * 1e-3,
            "samples_per_s": samples / (ms * 1e-3), "ms_per_iteration": ms / iters, "mrays_s_per_gpu": (prim + shad) / (iters + warm) * iters / (ms * 1e-3) / 1e6, "iterations_per_pass": per_pass,
            "framebuffer_allreduce_bytes": width * height * 12 if dist is not None else 0, "overlap": bool(loop.overlap), "mean_radiance": mean}


def measure_allreduce(dist, width=1920, height=1080, reps=20):
    """The collective alone: NCCL sum of one W x H x 3 float framebuffer, CUDA events on the stream it is enqueued on, max over ranks."""
    import torch
    if dist is None:
        return 0.0
    fb = torch.zeros(width * height * 3, dtype=torch.float32, device="cuda")
    comm = torch.cuda.Stream()
    with torch.cuda.stream(comm):
        for _ in range(3):
            dist.all_reduce(fb, op=dist.ReduceOp.SUM)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(comm)
        for _ in range(reps):
            dist.all_reduce(fb, op=dist.ReduceOp.SUM)
        e1.record(comm)
    comm.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def parity_sample(flat, scene, batches, stride=16):
    """After the timed regions: every `stride`-th ray of each bench batch re-traced by the CUDA path and by the CPU oracle (the
    checker, never the thing measured).  Closest hit: whole 16-byte records compared; occlusion: the hit / miss flag."""
    import torch
    from imbatracer_b200 import api
    from tests import refapi
    checked = mismatches = 0
    per_batch = {}
    for name, any_hit in BATCH_ORDER:
        r = np.ascontiguousarray(batches[name][::stride])
        got = api.hits_from_device(scene.trace(api.rays_to_device(r), any_hit=bool(any_hit)))
        torch.cuda.synchronize()
        want = refapi.trace_oracle(flat, r, any_hit=bool(any_hit))
        bad = int(((got["tri_id"] >= 0) != (want["tri_id"] >= 0)).sum()) if any_hit else int((got.view(np.uint8).reshape(-1, 16) != want.view(np.uint8).reshape(-1, 16)).any(1).sum())
        checked += len(r); mismatches += bad; per_batch[name] = bad
    return {"checked": checked, "mismatches": mismatches, "per_batch": per_batch, "stride": stride,
            "note": "CUDA path vs oracle/traversal_oracle.cpp on the same rays: closest hit bit-exact (tri_id, inst_id, t, u), occlusion flag-equal; the only known "
                    "exceptions are the 4 coplanar shadow rays of the full incoherent batch (tests/test_bench_parity_gpu.py)"}


def run_integrators(scene, flat, width=1920, height=1080, iters=4):
    """configs[3]: the bidirectional family at 1080p on the bench scene (rank 0, untimed region): light paths = W*H/2
    (test/run_bench.py:111), 1 connection, k = 10; ms per progressive iteration and rays per second."""
    from imbatracer_b200 import api
    out = {}
    for algo in ("bpt", "vcm", "lt", "ppm"):
        # one render(iters) call: BPT / VCM / PPM pipeline the iterations, the light tracer traces their light paths as one pass
        it = api.Integrator(scene, api.render_settings(flat, algo, width, height, seed=5, iterations_per_pass=iters if algo == "lt" else 1))
        it.preprocess(); it.render(iters if algo == "lt" else 1); it.sync()
        p0, s0 = it.ray_counts()
        t0 = time.perf_counter(); it.render(iters); it.sync(); dt = time.perf_counter() - t0
        p1, s1 = it.ray_counts()
        img = it.download()
        it.close()
        out[algo] = {"ms_per_iteration": dt / iters * 1e3, "samples_per_s": width * height * iters / dt, "mrays_s": (p1 - p0 + s1 - s0) / dt / 1e6,
                     "mean_radiance": float(img.mean()) / (iters + (iters if algo == "lt" else 1))}
    out["workload"] = f"{width}x{height}, light paths W*H/2, 1 connection, 10 nearest photons, max path length 10, {SCENE}"
    return out


def device_build_entry(flat, d_rays, hits, n):
    """Row f1 end to end: imba_scene_create with IMBA_ACCEL_DEVICE_BUILD (LBVH + 8-wide collapse + emission on the device,
    csrc/bvh8_gpu.cu) and closest-hit traversal of the same rays on the resulting tree."""
    import torch
    from imbatracer_b200 import api, procedural
    api.Scene(procedural.terrain(8), device_build=True).close()   # first use loads the builder's kernels (lazy module loading): not timed
    t0 = time.perf_counter(); sc = api.Scene(flat, device_build=True); wall = time.perf_counter() - t0
    info, st = sc.build_info(), sc.stats()
    cnt = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    sc.trace(d_rays, hits, counters=cnt); torch.cuda.synchronize()
    c = cnt.double().mean(0).cpu().numpy()
    for _ in range(3):
        sc.trace(d_rays, hits)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sc.trace(d_rays, hits)
    e1.record(); torch.cuda.synchronize()
    out = {"build_device_ms": info["device_seconds"] * 1e3, "scene_create_call_s": wall, "mesh_levels": info["mesh_levels"], "bvh8_bytes": int(st["bytes"]),
           "closest_mrays_s": n * 5 / e0.elapsed_time(e1) / 1e3, "nodes_per_ray": c[0], "tris_per_ray": c[1]}
    sc.close()
    return out


def run_extra_scenes(rank):
    """configs[4] and a memory-bound case (rank 0 only, after the timed regions): incoherent traversal on the procedural
    ~19 M-triangle instanced scene and on a 4.7 M-triangle unique terrain (292 MB of nodes + triangles, > L2), plus VCM with
    photon merging on the instanced scene.  Structures built by the product's host builder (csrc/accel_build.cpp)."""
    import torch
    from imbatracer_b200 import api, procedural
    from tests import raygen
    out = {}
    base = load_cornell("cornell_water")
    n = 1 << 22
    for label, make in (("cornell_water", lambda: base), ("instanced_19M", lambda: procedural.instanced_grid(base, 14, 14, 14)), ("terrain_4.7M", lambda: procedural.terrain(1536))):
        flat = make()
        scene = api.Scene(flat)
        rays = raygen.random_rays(flat, n, 11, tmax_mix=False); rays["tmin"] = 0
        d = api.rays_to_device(rays); hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        cnt = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
        scene.trace(d, hits, counters=cnt); torch.cuda.synchronize()
        c = cnt.double().mean(0).cpu().numpy()
        r = {"triangles_unique": int(flat.tri_count), "instances": int(len(flat.instances)), "bvh8_bytes": int(scene.stats()["bytes"])}
        for any_hit in (0, 1):
            for _ in range(3):
                scene.trace(d, hits, any_hit=bool(any_hit))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                scene.trace(d, hits, any_hit=bool(any_hit))
            e1.record(); torch.cuda.synchronize()
            r["occluded_mrays_s" if any_hit else "closest_mrays_s"] = n * 5 / e0.elapsed_time(e1) / 1e3
        bpr = 48 + 80 * c[0] + 48 * c[1] + 64 * c[2]
        r.update({"nodes_per_ray": c[0], "tris_per_ray": c[1], "instances_per_ray": c[2], "bytes_per_ray": bpr,
                  "algorithmic_gb_s": r["closest_mrays_s"] * 1e6 * bpr / 1e9, "rays": "4 Mi uniformly random rays in the bounding sphere"})
        if label.startswith("instanced"):
            integ = api.Integrator(scene, api.render_settings(flat, "vcm", 1280, 720, seed=3))
            integ.preprocess(); integ.render(1); integ.sync()
            t0 = time.perf_counter(); integ.render(4); integ.sync(); dt = time.perf_counter() - t0
            r["vcm_1280x720"] = {"ms_per_iteration": dt / 4 * 1e3, "samples_per_s": 1280 * 720 * 4 / dt, "light_paths": 1280 * 720 // 2, "connections": 1, "knn": 10}
            integ.close()
        if label.startswith("terrain"):  # row f1: the same mesh through the GPU builder (csrc/lbvh.cu)
            t0 = time.perf_counter(); procedural.build_accel(flat, gpu=True); wall = time.perf_counter() - t0
            scene2 = api.Scene(flat)
            for _ in range(3):
                scene2.trace(d, hits)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                scene2.trace(d, hits)
            e1.record(); torch.cuda.synchronize()
            r["gpu_lbvh"] = {"build_device_ms": flat.build_seconds * 1e3, "build_call_s": wall, "closest_mrays_s": n * 5 / e0.elapsed_time(e1) / 1e3}
            scene2.close()
            r["device_build"] = device_build_entry(flat, d, hits, n)
        out[label] = r
        scene.close()
    return out


def bind_to_gpu_numa_node(local):
    """One process per GPU: run (and first-touch the pinned host queues) on the cores of the GPU's own NUMA node, so that eight
    ranks do not stream their 48 bytes per ray across the socket interconnect.  Returns a description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, bits in enumerate(words) for b in range(64) if (bits >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cores of the GPU's NUMA node ({min(allowed)}-{max(allowed)})"
        return "GPU-local cores not in this process's cpuset: not bound"
    except Exception as e:   # no NVML / no permission: run unbound
        return f"not bound ({type(e).__name__})"


def host_copy_bandwidth(barrier, mib=256, reps=4):
    """Plain host memory copy bandwidth of this rank while all ranks copy at the same time (GB/s, read + write bytes): the ceiling
    the host-queue path shares between the ranks, beside the per-GPU PCIe rate."""
    a = np.ones(mib << 18, np.float32); b = np.empty_like(a)
    np.copyto(b, a)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        np.copyto(b, a)
    return 2 * a.nbytes * reps / (time.perf_counter() - t0) / 1e9


def raw_pcie_bandwidth(barrier, mib=256, reps=4):
    """What the host can feed: plain cudaMemcpyAsync of pinned buffers, host->device and device->host at the same time on two
    streams, all ranks at once (GB/s of this rank, both directions added).  The host-queue path cannot beat the sum over ranks."""
    import torch
    n = mib << 18
    h_up = torch.empty(n, dtype=torch.float32).pin_memory(); h_dn = torch.empty(n // 2, dtype=torch.float32).pin_memory()   # 2 : 1 like rays : hits
    d_up = torch.empty(n, dtype=torch.float32, device="cuda"); d_dn = torch.empty(n // 2, dtype=torch.float32, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    def go():
        with torch.cuda.stream(s_up):
            d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_dn.copy_(d_dn, non_blocking=True)
    go(); torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    return (h_up.nbytes + h_dn.nbytes) * reps / (time.perf_counter() - t0) / 1e9


def run_ours(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    binding = bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get("IMBA_BENCH_NO_BIND") else "single rank: not bound"
    import torch
    import torch.distributed as dist
    from imbatracer_b200 import abi, api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = abi.lib()
    flat = load_flat()
    scene = api.Scene(flat)
    n = RAYS_PER_BATCH
    batches = make_batches(flat, n, 0x1234ABCD + rank, scene=scene)
    dev = {k: api.rays_to_device(v) for k, v in batches.items()}
    hits = {k: torch.empty((n, 4), dtype=torch.int32, device="cuda") for k in batches}
    pinned_rays = {k: torch.from_numpy(v.view(np.float32).reshape(-1, 8)).pin_memory() for k, v in batches.items()}
    pinned_hits = {k: torch.empty((n, 4), dtype=torch.int32).pin_memory() for k in batches}
    stream = torch.cuda.current_stream()

    def step(events=None):
        for i, (name, any_hit) in enumerate(BATCH_ORDER):
            if events is not None:
                events[i][0].record(stream)
            scene.trace(dev[name], hits[name], any_hit=bool(any_hit), count=n)
            if events is not None:
                events[i][1].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = L.imba_launch_count()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in BATCH_ORDER] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step(ev[k])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.imba_launch_count() - launches0
    per_kernel_ms = [float(np.mean([a.elapsed_time(b) for a, b in [ev[k][i] for k in range(args.steps)]])) for i in range(len(BATCH_ORDER))]

    # end to end through the host-buffer call (pinned host queues; H2D + D2H inside the timed region)
    # the four queues are traversed by four host threads, as the reference's TileScheduler threads do (tile_scheduler.h:69-78):
    # every call is synchronous for its caller, the library overlaps the calls (one staging pipeline per concurrent caller)
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(len(BATCH_ORDER))

    def e2e_one(item):
        name, any_hit = item
        abi.check(L.imba_trace_host(scene.h, pinned_rays[name].data_ptr(), pinned_hits[name].data_ptr(), n, any_hit), "imba_trace_host")

    def e2e_step():
        list(pool.map(e2e_one, BATCH_ORDER))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    host_bw = host_copy_bandwidth(barrier) if world > 1 else 0.0
    raw_pcie = raw_pcie_bandwidth(barrier)
    # configs[2] AS STATED: PT 1920x1080, 256 spp IN TOTAL split over the ranks (strong scaling), one NCCL sum of the float framebuffer
    # per step.  8 progressive iterations per wavefront pass (imba_render_settings::iterations_per_pass) and per sum; every path keeps
    # the seed of its own iteration, the image is that of single passes -- which `pt_single_pass` times for comparison
    d = dist if world > 1 else None
    pt = run_pt(scene, flat, rank, world, barrier, d, iters=256 // world, per_pass=8)
    pt_serial = run_pt(scene, flat, rank, world, barrier, d, iters=256 // world, per_pass=8, overlap=False) if world > 1 else pt
    pt_single = run_pt(scene, flat, rank, world, barrier, d, iters=max(4, min(args.steps * 2, 32)))
    # the same with 4 samples per pixel per wavefront pass: larger wavefronts amortise the tails of the late, small waves
    pt_spp4 = run_pt(scene, flat, rank, world, barrier, d, iters=max(2, min(args.steps, 8)), spp=4)
    allreduce_ms = measure_allreduce(d)
    sampler.stop_flag = True; sampler.join(timeout=2)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    bw = torch.tensor([host_bw, raw_pcie], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(bw, op=dist.ReduceOp.SUM)
    ms, e2e_ms = float(t[0]), float(t[1])
    total_rays = 4 * n * args.steps * world
    value = total_rays / (ms * 1e-3) / 1e6
    e2e_value = 4 * n * e2e_steps * world / (e2e_ms * 1e-3) / 1e6

    line = None
    if rank == 0:
        # algorithmic bytes per ray of the dominant kernel (incoherent closest-hit): B = 32 + 16 + N_node*80 + N_tri*48
        #  with N_node / N_tri = per-ray mean fetch counts of the instrumented kernel on the same batch
        cnt = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
        scene.trace(dev["incoherent_closest"], hits["incoherent_closest"], counters=cnt, count=n)
        torch.cuda.synchronize()
        c = cnt.double().mean(0).cpu().numpy()
        bytes_per_ray = 32 + 16 + c[0] * 80 + c[1] * 48 + c[2] * 64
        peak, peak_src = measured_peak_hbm()
        dom_ms = per_kernel_ms[1]
        achieved = n * bytes_per_ray / (dom_ms * 1e-3) / 1e9
        cores = os.cpu_count()
        sample = 1 << 20
        cpu_total, cpu_secs = cpu_arm(flat, batches, sample, cores)
        port = {"value": cpu_total / cpu_secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                "sample": f"first {sample} rays of each of the 4 batches, oracle/traversal_oracle.cpp (scalar restatement of the four entry points) with {cores} threads, "
                          f"on the same BVH2 structure the CUDA path converted"}
        # the record's CPU baseline = the reference's OWN host pipeline (N = 1 only: its loader, SBVH builder, TileScheduler, RayQueue and
        # PathTracer, unmodified, from oracle/_ref) on BASELINE.json configs[0]; the traversal inside it is the scalar restatement,
        # because the AnyDSL kernels are not in the reference tree.  The traversal-only figure stays beside it.
        ref0 = reference_cfg0(cores) if world == 1 else {"unavailable": "timed at N = 1 only"}
        if "mrays_s" in ref0:
            cpu_baseline = {"value": ref0["mrays_s"], "unit": "Mrays/s", "cores": cores, "kind": "reference",
                            "sample": ref0["workload"] + f"; {ref0['seconds']:.2f} s, {ref0['samples_per_s'] / 1e6:.2f} M samples/s; unmodified reference host pipeline (TileScheduler + PathTracer, "
                                      "oracle/_ref), its traversal entry points = the scalar restatement oracle/traversal_oracle.cpp",
                            "traversal_only": port}
        else:
            cpu_baseline = dict(port, reference_pipeline=ref0)
        device_build = device_build_entry(flat, dev["incoherent_closest"], hits["incoherent_closest"], n)
        device_build["rays"] = "the incoherent closest-hit batch"
        traffic, traffic_src = ncu_traffic()
        config = workload_config(flat, n)
        config.update({"pt_workload": "configs[2]: PT 1920x1080, 256 spp in total split over the ranks (strong scaling), max path length 10, 8 iterations per wavefront pass and per "
                                      "framebuffer sum, on " + SCENE,
                       "pt_total_spp": pt["total_spp"], "pt_seconds": pt["seconds"], "pt_samples_per_s": pt["samples_per_s"], "pt_ms_per_iteration": pt["ms_per_iteration"],
                       "pt_ms_per_iteration_no_overlap": pt_serial["ms_per_iteration"], "pt_samples_per_s_no_overlap": pt_serial["samples_per_s"],
                       "allreduce_bytes": pt["framebuffer_allreduce_bytes"], "allreduce_ms": allreduce_ms,
                       "allreduce_note": "NCCL sum of the 1920x1080x3 float framebuffer, timed alone on its stream (max over ranks); in the PT loop it runs once per 8 iterations, "
                                         "overlapped with the next pass (pt_ms_per_iteration) or not (pt_ms_per_iteration_no_overlap)",
                       "pt_single_pass_samples_per_s": pt_single["samples_per_s"], "pt_mean_radiance": pt["mean_radiance"]})
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 4 * n * 32, "d2h_bytes_per_step": 4 * n * 16,
                        "steps": e2e_steps, "note": "imba_trace_host on pinned host queues from 4 host threads (one per queue), wall clock",
                        "pcie_gb_s_per_gpu": e2e_value * 1e6 * 48 / 1e9 / world, "host_cores": os.cpu_count(), "cpu_binding_rank0": binding,
                        "host_copy_gb_s_all_ranks": float(bw[0]) if world > 1 else None,
                        "raw_memcpy_gb_s_all_ranks": float(bw[1]),
                        "limit_note": "pcie_gb_s_per_gpu x n_gpus against raw_memcpy_gb_s_all_ranks (pinned cudaMemcpyAsync up and down at once on every rank, no kernels): "
                                      "what the host's memory / PCIe complex delivers to all GPUs together bounds this path, not the traversal"},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "parity_sample": parity_sample(flat, scene, batches),
                "cfg0_reference_on_library": reference_on_library_cfg0(cores) if world == 1 else {"unavailable": "timed at N = 1 only"},
                "pt": pt, "pt_no_overlap": pt_serial, "pt_single_pass": pt_single, "pt_spp4": pt_spp4, "cfg0": ours_cfg0(), "integrators": run_integrators(scene, flat), "scenes": run_extra_scenes(rank), "device_build": device_build,
                "per_kernel": {name: {"ms": per_kernel_ms[i], "mrays_s": n / per_kernel_ms[i] / 1e3} for i, (name, _) in enumerate(BATCH_ORDER)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "kernel": "trace_kernel<closest> on the incoherent batch", "peak_source": peak_src,
                             "bytes_per_ray": bytes_per_ray, "nodes_per_ray": c[0], "tris_per_ray": c[1], "instances_per_ray": c[2],
                             "algorithmic_bytes_per_launch": bytes_per_ray * n, "structure_bytes": int(scene.stats()["bytes"]),
                             "note": "algorithmic bytes = ray 32 + hit 16 + 80/node + 48/triangle + 64/instance fetched (counted by the instrumented kernel on "
                                     "this batch); traffic = dram bytes of the same launch from the committed ncu capture (profiles/ncu_traffic.json, refused unless it was "
                                     "taken on the same kernel sources); the difference is what L1/L2 serve -- the kernel is issue-bound, not HBM-bound (DESIGN.md section 3)"},
                "cpu_baseline": cpu_baseline}
    scene.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if line is not None:
        emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (NCCL prints its version line with printf) must not land on stdout beside the ONE JSON line: everything
    written to file descriptor 1 while the bench runs goes to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
