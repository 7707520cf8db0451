import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SCENE_NAMES = ["cornell_org", "cornell_indirect", "cornell_water", "cornell_instances", "cornell_specular_front"]


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
