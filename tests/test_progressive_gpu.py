"""ProgressiveRenderer (imbatracer_b200/distributed.py): the two-framebuffer loop that takes the per-iteration sum off the
critical path must accumulate exactly the same iterations as the plain loop -- also for VCM, whose light pass replays CUDA
graphs that captured the framebuffer pointer (csrc/integrator.cu: imba_integrator_set_framebuffer parks one graph per buffer)."""
import numpy as np
import pytest

from tests.conftest import load_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["pt", "vcm", "lt"])
def test_overlapped_accumulation_equals_plain_loop(algo):
    import torch
    from imbatracer_b200 import api
    from imbatracer_b200.distributed import ProgressiveRenderer
    flat = load_scene("cornell_water", "gpu")     # glass: long light paths, so the VCM tail graph is used
    sc = api.Scene(flat)
    imgs = []
    for overlap in (False, True):
        it = api.Integrator(sc, api.render_settings(flat, algo, 96, 96, seed=11))
        it.preprocess()
        loop = ProgressiveRenderer(it, 96, 96, overlap=overlap)
        loop.step(7)
        img = loop.image().cpu().numpy()
        assert loop.iterations == 7 and np.isfinite(img).all() and img.sum() > 0
        imgs.append(img)
        it.close()
    sc.close()
    a, b = imgs
    # same seeds, same iterations: only the order of the atomic float adds inside an iteration differs
    assert abs(a.sum() - b.sum()) <= 2e-4 * a.sum()
    assert np.abs(a - b).max() <= 1e-3 * max(1.0, a.max())
