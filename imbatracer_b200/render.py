"""Headless driver around the device integrators with the reference's command line (frontend/cmd_line.h:132-242) and its
final report line (frontend/render_window.cpp:84-88), so that test/run_bench.py-style tooling keeps working:

  python -m imbatracer_b200.render scene.npz [out.png] -a pt|bpt|vcm|ppm|lt|vcm_pt -w 512 -h 512 -s SAMPLES | -t SECONDS
         [--spp N] [--max-path-len N] [--light-path-count N] [-r RADIUS_FACTOR] [-c CONNECTIONS] [-k KNN] [--gamma G]
         [--compare ref.png]

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
    a = ap.parse_args(argv)
    if a.s <= 0 and a.t <= 0:
        a.s = 16

    import torch
    import torch.distributed as dist
    from . import api
    from .distributed import ProgressiveRenderer
    from .scene import FlatScene
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    flat = FlatScene.load(a.scene)
    scene = api.Scene(flat)
    st = api.render_settings(flat, a.algo, a.w, a.h, fov=a.f, spp=a.spp, max_path_len=a.max_path_len, light_path_count=a.light_path_count,
                             num_connections=a.c, num_knn=a.k, radius_factor=a.r, seed=a.seed, rank=rank, rank_count=world)
    integ = api.Integrator(scene, st)
    integ.preprocess()
    loop = ProgressiveRenderer(integ, a.w, a.h, a.spp)
    t0 = time.time()
    while True:
        loop.step(1)
        samples = loop.iterations * world * a.spp
        if a.s > 0 and samples >= a.s:
            break
        if a.t > 0:
            integ.sync()
            if time.time() - t0 >= a.t:
                break
    img = loop.image().cpu().numpy()
    secs = time.time() - t0
    if rank == 0:
        frames = loop.iterations * world
        print(f"Done after {secs} seconds, {frames * a.spp} samples @ {frames / secs} frames per second, {1000.0 * secs / frames}ms per frame")
        prim, shad = integ.ray_counts()
        print(f"{prim} primary ray(s), {shad} shadow ray(s) on rank 0")
        png = to_png_array(img, a.gamma)
        if a.out:
            from PIL import Image
            Image.fromarray(png).save(a.out)
        if a.compare:
            from PIL import Image
            ref = np.asarray(Image.open(a.compare).convert("RGB").resize((a.w, a.h), Image.BOX))
            print(f"RMSE {rmse_like_imagemagick(png, ref)}")
    integ.close(); scene.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
