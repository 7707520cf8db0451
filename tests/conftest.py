import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SCENE_NAMES = ["cornell_org", "cornell_indirect", "cornell_water", "cornell_instances", "cornell_specular_front", "stilllife_proxy"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def built():
    """Builds the product library and the oracle if they are missing (both are git-ignored build outputs)."""
    import __graft_entry__ as g
    g.build()


def load_scene(name, layout):
    from imbatracer_b200.scene import FlatScene
    return FlatScene.load(os.path.join(GOLDEN, f"scene_{name}_{layout}.npz"), name)


def load_hits(name):
    return np.load(os.path.join(GOLDEN, f"hits_{name}.npz"))


def batch_for(name, flat):
    from tests import raygen
    return raygen.test_batch(flat, load_hits(name)["camera_rays"])


def zero_distance_ties(rays, *results):
    """Documented near-tie class: a ray with tmin = 0 that STARTS on geometry hits it at t = 0; when it
    also leaves the bounds through that very point (a room corner), whether any traversal reaches the triangle depends on how
    a box touched in one point is culled.  Returns the mask of rays for which some result reports such a 0-distance hit."""
    m = np.zeros(len(rays), bool)
    for h in results:
        m |= (h["tri_id"] >= 0) & (h["tmax"] == 0.0)
    return m & (rays["tmin"] == 0.0)


def assert_hits_equal(rays, got, want, max_tie_fraction=1e-3):
    """Bit-exact equality of hit records, except inside the 0-distance tie class (bounded in size)."""
    if got.tobytes() == want.tobytes():
        return 0
    ties = zero_distance_ties(rays, got, want)
    differ = (got.view(np.uint8).reshape(-1, 16) != want.view(np.uint8).reshape(-1, 16)).any(1)
    bad = np.nonzero(differ & ~ties)[0]
    assert len(bad) == 0, (len(bad), got[bad[:3]], want[bad[:3]], rays[bad[:3]])
    assert differ.sum() <= max(1, int(max_tie_fraction * len(rays))), int(differ.sum())
    return int(differ.sum())
