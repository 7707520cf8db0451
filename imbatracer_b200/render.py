"""Headless driver around the device integrators with the reference's command line (frontend/cmd_line.h:132-242) and its
final report line (frontend/render_window.cpp:84-88), so that test/run_bench.py-style tooling keeps working:

  python -m imbatracer_b200.render scene.npz [out.png] -a pt|bpt|vcm|ppm|lt|vcm_pt -w 512 -h 512 -s SAMPLES | -t SECONDS
         [--spp N] [--max-path-len N] [--light-path-count N] [-r RADIUS_FACTOR] [-c CONNECTIONS] [-k KNN] [--gamma G]
         [--compare ref.png] [--iterations-per-pass N]

`scene.npz` is a flattened scene (imbatracer_b200/scene.py), i.e. what the reference's loader + BVH builders produce and a
host program hands to imba_scene_create; loading OBJ/MTL files is the reference's job (out of scope).
The PNG is written like loaders/store_png.h:59-62 (8 bit, value = clamp(radiance ** gamma)); --compare prints the RMSE the
reference's benchmark script computes with ImageMagick (test/run_bench.py:113-129), here with numpy.  Under torchrun every
rank renders its share of the progressive iterations and the framebuffers are summed with NCCL once per iteration."""
import argparse
import os
import sys
import time

import numpy as np


def to_png_array(radiance, gamma=0.5):
    return (255.0 * np.clip(np.power(np.clip(radiance, 0, None), gamma), 0.0, 1.0)).astype(np.uint8)


def rmse_like_imagemagick(a8, b8):
    """compare -metric RMSE: root of the mean squared difference over all channels, normalised to [0, 1]."""
    d = a8.astype(np.float64) / 255.0 - b8.astype(np.float64) / 255.0
    return float(np.sqrt(np.mean(d * d)))


def compare_with_reference(radiance, ref_path, gamma=0.5):
    """RMSE of the image against a reference: a PNG (box-resized to the image size) or one of the committed
    tests/golden/ref_*_128.npy files (linear radiance, 128x128; the image is box-filtered down to it first)."""
    h, w, _ = radiance.shape
    if ref_path.endswith(".npy"):
        ref = np.load(ref_path)
        rh, rw, _ = ref.shape
        small = radiance[:h - h % rh, :w - w % rw].reshape(rh, h // rh, rw, w // rw, 3).mean(axis=(1, 3))
        return rmse_like_imagemagick(to_png_array(small, gamma), to_png_array(ref, gamma))
    from PIL import Image
    ref = np.asarray(Image.open(ref_path).convert("RGB").resize((w, h), Image.BOX))
    return rmse_like_imagemagick(to_png_array(radiance, gamma), ref)


def main(argv=None):
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("scene"); ap.add_argument("out", nargs="?")
    ap.add_argument("-a", dest="algo", default="pt"); ap.add_argument("-w", type=int, default=512); ap.add_argument("-h", type=int, default=512)
    ap.add_argument("-s", type=int, default=0, help="max samples per pixel"); ap.add_argument("-t", type=float, default=0.0, help="max seconds")
    ap.add_argument("-q", action="store_true"); ap.add_argument("-f", type=float, default=60.0)
    ap.add_argument("--spp", type=int, default=1); ap.add_argument("--max-path-len", type=int, default=10)
    ap.add_argument("--light-path-count", type=int, default=0); ap.add_argument("-r", type=float, default=2.0)
    ap.add_argument("-c", type=int, default=1); ap.add_argument("-k", type=int, default=10); ap.add_argument("--gamma", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--compare"); ap.add_argument("--help", action="help")
    ap.add_argument("--intermediate-path"); ap.add_argument("--intermediate-time", type=float, default=10.0)
    ap.add_argument("--iterations-per-pass", type=int, default=1, help="progressive iterations per step: PT renders them as one wavefront pass, the bidirectional family pipelines them (same image)")
    a = ap.parse_args(argv)
    if a.s <= 0 and a.t <= 0:
        a.s = 16

    import torch
    import torch.distributed as dist
    from . import api
    from .distributed import ProgressiveRenderer, agree_to_stop
    from .scene import FlatScene
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    flat = FlatScene.load(a.scene)
    scene = api.Scene(flat)
    st = api.render_settings(flat, a.algo, a.w, a.h, fov=a.f, spp=a.spp, max_path_len=a.max_path_len, light_path_count=a.light_path_count,
                             num_connections=a.c, num_knn=a.k, radius_factor=a.r, seed=a.seed, rank=rank, rank_count=world,
                             iterations_per_pass=a.iterations_per_pass)
    integ = api.Integrator(scene, st)
    integ.preprocess()
    loop = ProgressiveRenderer(integ, a.w, a.h, a.spp, iterations_per_step=a.iterations_per_pass)   # PT: one wavefront pass; BPT / VCM / PPM: pipelined iterations
    t0 = time.time()
    next_dump = a.intermediate_time
    while True:
        loop.step(loop.per_step)
        samples = loop.iterations * world * a.spp
        if a.s > 0 and samples >= a.s:
            break
        if a.t > 0 or a.intermediate_path:
            integ.sync()
            el = time.time() - t0
            if a.intermediate_path and el >= next_dump:   # convergence series, render_window.cpp:72-81
                part = loop.image().cpu().numpy()
                if rank == 0:
                    from PIL import Image
                    Image.fromarray(to_png_array(part, a.gamma)).save(f"{a.intermediate_path}{int(el * 1000)}ms.png")
                next_dump += a.intermediate_time
            if a.t > 0 and agree_to_stop(el >= a.t):   # collective: all ranks leave after the same step
                break
    img = loop.image().cpu().numpy()
    secs = time.time() - t0
    if rank == 0:
        frames = loop.iterations * world
        png = to_png_array(img, a.gamma)
        if a.out:
            from PIL import Image
            Image.fromarray(png).save(a.out)
        if a.compare:
            print(f"RMSE {compare_with_reference(img, a.compare, a.gamma)}")
        # the two lines test/run_bench.py parses (its regexes at :170 and :203), last in the output
        print(f"Done after {secs} seconds, {frames * a.spp} samples @ {frames / secs} frames per second, {1000.0 * secs / frames}ms per frame")
        prim, shad = integ.ray_counts()
        print(f"Number primary rays: {prim * world} Number shadow rays: {shad * world}")
    integ.close(); scene.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
