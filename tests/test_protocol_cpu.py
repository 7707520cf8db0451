"""Host-side pieces of the benchmark protocol (tools/run_bench.py, imbatracer_b200/render.py): output parsing as in the
reference's test/run_bench.py:160-216 and the numpy replacement of its ImageMagick RMSE (:113-129)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_parse_reference_report_lines():
    import run_bench
    text = ("RMSE 0.0123\nDone after 30.5 seconds, 64 samples @ 2.09836 frames per second, 476.562ms per frame\n"
            "Number primary rays: 1000 Number shadow rays: 500\n")
    r = run_bench.parse_output(text)
    assert r["time"] == 30.5 and r["samples"] == 64 and r["rays"] == 1500 and abs(r["rmse"] - 0.0123) < 1e-12
    assert abs(r["ms_per_frame"] - 476.562) < 1e-9
    # the original accepts a bare "Done" line as the last line too
    assert run_bench.parse_output("Done after 1 seconds, 2 samples @ 2 frames per second, 500ms per frame\n")["rays"] == 0
    assert run_bench.parse_output("Segmentation fault\n") is None


def test_rmse_matches_definition_and_npy_reference():
    from imbatracer_b200 import render
    a = np.zeros((4, 4, 3), np.uint8); b = np.full((4, 4, 3), 255, np.uint8)
    assert render.rmse_like_imagemagick(a, a) == 0.0 and abs(render.rmse_like_imagemagick(a, b) - 1.0) < 1e-12
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_cornell_org_128.npy"))
    up = np.repeat(np.repeat(ref, 4, axis=0), 4, axis=1)          # a 512x512 image whose box filter is the reference
    assert render.compare_with_reference(up, os.path.join(ROOT, "tests", "golden", "ref_cornell_org_128.npy")) < 2e-3  # 8-bit quantisation
    assert render.compare_with_reference(up * 0.5, os.path.join(ROOT, "tests", "golden", "ref_cornell_org_128.npy")) > 0.05


def test_bench_settings_point_at_committed_fixtures():
    import run_bench
    for s in run_bench.BENCH_SETTINGS:
        assert os.path.exists(os.path.join(run_bench.GOLD, s["scene"])), s["scene"]
        assert os.path.exists(os.path.join(run_bench.GOLD, s["reference"])), s["reference"]
