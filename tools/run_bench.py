"""Python-3 restatement of the authors' benchmark protocol (reference test/run_bench.py:58-300) on top of the headless
driver `python -m imbatracer_b200.render`:

  python tools/run_bench.py [-t SECONDS] [-a pt,bpt,vcm,...] [--gpus N] [--convergence] [--out results/]

For every (scene, algorithm, scheduling setting) it renders for a fixed time, parses the driver's "Done after ..." and
"Number primary rays ..." lines exactly like the original script, computes the RMSE against the converged reference and
writes `results/bench_<date>.csv` with the original columns
  name, algorithm, time, samples, fps, ms_per_frame, rmse, rays_per_sec, scheduling
The reference images are the committed tests/golden/ref_*_128.npy (box-filtered copies of test/references/*.png); the
scenes are the flattened Cornell fixtures (the Still Life / Sponza / car geometry is not in the reference tree).
ImageMagick's `compare -metric RMSE` is replaced by numpy (imbatracer_b200/render.py:rmse_like_imagemagick)."""
import argparse
import datetime
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# run_bench.py:7-88 -- the entries whose geometry exists here; same names, sizes and extra arguments
BENCH_SETTINGS = [
    {"name": "Cornell box", "scene": "scene_cornell_org_gpu.npz", "reference": "ref_cornell_org_128.npy", "width": 1024, "height": 1024,
     "base_filename": "cornell", "args": []},
    {"name": "Cornell specular balls", "scene": "scene_cornell_specular_front_gpu.npz", "reference": "ref_cornell_specular_front_128.npy",
     "width": 1024, "height": 1024, "base_filename": "cornell_specular_front", "args": []},
    {"name": "Cornell indirect", "scene": "scene_cornell_indirect_gpu.npz", "reference": "ref_cornell_indirect_128.npy", "width": 1024,
     "height": 1024, "base_filename": "cornell_indirect", "args": []},
    {"name": "Cornell water", "scene": "scene_cornell_water_gpu.npz", "reference": "ref_cornell_water_128.npy", "width": 1024, "height": 1024,
     "base_filename": "cornell_water", "args": []},
]
LIGHT_PATH_FRAC = 0.5  # run_bench.py:111


def scheduler_args(samples=(1,), connections=(1,)):
    out = []
    for s in samples:                      # run_bench.py:96-106 (thread count and tile size have no device-side meaning)
        for c in connections:
            out.append({"name": f"samples: {s}, connections: {c}", "abbr": f"{s}{c}", "args": ["-c", str(c), "--spp", str(s)], "samples_per_frame": s})
    return out


def parse_output(text):
    """run_bench.py:160-216: the last line starting with D is the performance line, N lines after it are ray counts."""
    lines = [l for l in text.splitlines() if l.strip()]
    perf, rays = None, 0
    for l in lines[::-1]:
        m = re.match(r"Number primary rays: (\d+) Number shadow rays: (\d+)", l)
        if m:
            rays += int(m.group(1)) + int(m.group(2))
            continue
        if l.startswith("D"):
            perf = l
        break
    if perf is None:
        return None
    m = re.match(r"Done after (\d+\.?\d*) seconds, (\d+) samples @ (\d+\.?\d*) frames per second, (\d+\.?\d*)ms per frame", perf)
    if m is None:
        return None
    rm = re.search(r"^RMSE (\S+)", text, re.M)
    return {"time": float(m.group(1)), "samples": int(m.group(2)), "fps": float(m.group(3)), "ms_per_frame": float(m.group(4)),
            "rays": rays, "rmse": float(rm.group(1)) if rm else float("nan")}


def run_one(setting, alg, sched, seconds, out_dir, gpus, convergence, step_sec):
    out_png = os.path.join(out_dir, f"{setting['base_filename']}_{alg}{sched['abbr']}_{seconds}sec.png")
    cmd = [sys.executable]
    if gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1", "--master-port", "29533",
                "-m", "imbatracer_b200.render"]
    else:
        cmd += ["-m", "imbatracer_b200.render"]
    cmd += [os.path.join(GOLD, setting["scene"]), out_png, "-w", str(setting["width"]), "-h", str(setting["height"]), "-q", "-t", str(seconds),
            "-a", alg, "--light-path-count", str(int(LIGHT_PATH_FRAC * setting["width"] * setting["height"])),
            "--compare", os.path.join(GOLD, setting["reference"])] + setting["args"] + sched["args"]
    if convergence:
        cmd += ["--intermediate-path", out_png[:-4] + "_conv_", "--intermediate-time", str(step_sec)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    r = parse_output(p.stdout)
    if r is None:
        print("render failed:\n" + p.stdout[-2000:] + p.stderr[-2000:])
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-t", type=int, default=30, help="seconds per run (run_bench.py:108: 30)")
    ap.add_argument("-a", default="vcm", help="comma separated algorithms (run_bench.py:109: vcm)")
    ap.add_argument("--gpus", type=int, default=1); ap.add_argument("--convergence", action="store_true")
    ap.add_argument("--step", type=int, default=5); ap.add_argument("--out", default=os.path.join(ROOT, "results"))
    ap.add_argument("--scenes", default="", help="comma separated base_filenames (default: all)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    rows = ["name,algorithm,time,samples,fps,ms_per_frame,rmse,rays_per_sec,scheduling"]
    for setting in BENCH_SETTINGS:
        if a.scenes and setting["base_filename"] not in a.scenes.split(","):
            continue
        print("== Running benchmark for " + setting["name"])
        for alg in a.a.split(","):
            for sched in scheduler_args():
                r = run_one(setting, alg, sched, a.t, a.out, a.gpus, a.convergence, a.step)
                if r is None:
                    continue
                rps = r["rays"] / r["time"] if r["rays"] else 0
                print(f"   > {alg}: {r['samples']} samples, {r['ms_per_frame']:.3f} ms per frame, RMSE {r['rmse']:.5f}, rays per second {rps:.4g}")
                rows.append(f"{setting['name']},{alg},{r['time']},{r['samples']},{r['fps']},{r['ms_per_frame']},{r['rmse']},{rps},{sched['name']}")
    path = os.path.join(a.out, "bench_" + datetime.datetime.now().strftime("%Y%m%d_%H%M%S") + ".csv")
    with open(path, "w") as f:
        f.write("\n".join(rows) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
