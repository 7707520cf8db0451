"""The procedural Still Life stand-in (imbatracer_b200/procedural.py) against the reference's loader: the committed fixture
scene_stilllife_proxy_*.npz is what the reference's OBJ/MTL/TGA loaders (loaders/load_obj.cpp, load_tga.cpp), convert_materials
and create_mesh (frontend/build_scene.cpp:55-302, 616-626) made of the files tests/obj_export.py wrote for the same generator
call; the arrays the generator hands to imba_scene_create directly must be the same scene."""
import numpy as np

from imbatracer_b200 import procedural
from tests import refapi
from tests.conftest import load_scene


def _corners(flat):
    a = flat.arrays
    idx = a["indices"].reshape(-1, 4)
    return (a["vertices"].reshape(-1, 4)[:, :3][idx[:, :3]], a["normals"].reshape(-1, 4)[:, :3][idx[:, :3]],
            a["texcoords"].reshape(-1, 2)[idx[:, :3]], idx[:, 3], a["geom_normals"].reshape(-1, 4))


def test_generator_equals_reference_loaded_scene():
    mine = procedural.still_life_proxy(refapi.PROXY_DETAIL, accel=None)
    ref = load_scene("stilllife_proxy", "gpu")
    assert mine.tri_count == ref.tri_count
    for x, y, what in zip(_corners(mine), _corners(ref), ("positions", "normals", "texcoords", "material ids", "geometric normals")):
        assert np.array_equal(x, y), what            # vertex ORDER differs (the loader renumbers), per-corner data may not
    for k in ("materials", "lights", "textures"):
        assert np.array_equal(mine.arrays[k].view(np.uint8), ref.arrays[k].view(np.uint8)), k
    assert np.array_equal(mine.texels, ref.texels)
    assert np.array_equal(mine.cam, ref.cam) and np.allclose(mine.bsphere, ref.bsphere, rtol=1e-6)
    # the 23 materials of still_life.mtl + the dummy + one duplicate per emissive triangle (build_scene.cpp:221-241)
    assert len(procedural.STILL_LIFE_MTL) == 23 and len(mine.materials) == 24 + len(mine.lights)
    kinds = {int(k) for k in mine.materials["kind"]}
    assert kinds == {0, 1, 2, 3}


def test_detail_scales_triangle_count():
    small = procedural.still_life_proxy(0.005, accel=None)
    big = procedural.still_life_proxy(0.05, accel=None)
    assert small.tri_count < big.tri_count and len(small.lights) == len(big.lights) == 24
    # no degenerate triangles survive (their geometric normal would be NaN)
    assert np.isfinite(big.geom_normals).all() and np.isfinite(big.normals).all()
