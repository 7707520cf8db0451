"""CPU tests of the host logic: the MBVH4/BVH2 -> compressed 8-wide converter (imbatracer_b200/csrc/bvh8_build.cpp),
checked structurally and, through the CPU walker in oracle/bvh8_walk.cpp, against the oracle."""
import numpy as np
import pytest

from imbatracer_b200 import abi
from tests import refapi
from tests.conftest import SCENE_NAMES, assert_hits_equal, batch_for, load_scene, zero_distance_ties


@pytest.mark.parametrize("name", SCENE_NAMES)
@pytest.mark.parametrize("layout", ["cpu", "gpu"])
def test_walker_on_converted_bvh_matches_oracle(name, layout):
    flat = load_scene(name, layout)
    rays = batch_for(name, flat)
    b8 = refapi.Bvh8Host(flat)
    want = refapi.trace_oracle(flat, rays)
    got, st = b8.walk(rays, stats=True)
    assert_hits_equal(rays, got, want)
    occ_w = refapi.trace_oracle(flat, rays, any_hit=True)
    occ_g = b8.walk(rays, any_hit=True)
    ties = zero_distance_ties(rays, got, want, occ_w, occ_g)
    assert np.array_equal((occ_w["tri_id"] >= 0)[~ties], (occ_g["tri_id"] >= 0)[~ties])
    assert ((occ_w["tri_id"] >= 0) != (occ_g["tri_id"] >= 0)).sum() <= len(rays) // 1000
    assert st[:, 3].max() + 6 <= 40  # traversal stack bound used by the kernel (kernels.h)


@pytest.mark.parametrize("name", ["cornell_water", "cornell_instances"])
@pytest.mark.parametrize("layout", ["cpu", "gpu"])
def test_converted_structure_invariants(name, layout):
    flat = load_scene(name, layout)
    b8 = refapi.Bvh8Host(flat)
    n = b8.nodes
    # every triangle id of the scene is present at least once; ids stay global
    ids = b8.tris["id"][:-1]
    assert set(ids.tolist()) == set(range(flat.tri_count))
    # the record carries the first two components of n = e1 x e2, computed unfused in float32 (bvh8.h; tri_test reads them)
    t = b8.tris[:-1]; e1, e2 = t["e1"], t["e2"]
    assert np.array_equal(t["nx"].view(np.uint32), (e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]).view(np.uint32))
    assert np.array_equal(t["ny"].view(np.uint32), (e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]).view(np.uint32))
    # meta encoding: inner children have 0b001_11sss with sss = slot and imask bit set; leaves unary counts 1/3/7
    for s in range(8):
        m = n["meta"][:, s].astype(np.int32)
        inner = (m & 0x18) == 0x18
        assert np.all((m[inner] >> 5) == 1) and np.all((m[inner] & 7) == s)
        assert np.all(((n["imask"] >> s) & 1).astype(bool) == inner)
        leaf = (~inner) & (m != 0)
        assert np.all(np.isin(m[leaf] >> 5, [1, 3, 7])) and np.all((m[leaf] & 31) < 24)
    # child ranges are in bounds
    n_inner = np.array([bin(int(x)).count("1") for x in n["imask"]])
    assert np.all(n["child_base"] + n_inner <= len(n))
    # instance records point at node indices and carry the reference's instance ids
    assert np.all(b8.insts["root"] < len(n)) and set(b8.insts["inst_id"].tolist()) == set(range(len(flat.instances)))
    # quantised boxes of used slots are non-empty; unused slots are inverted
    used = n["meta"] != 0
    assert np.all((n["q"][:, 0:3, :] <= n["q"][:, 3:6, :]).all(axis=1)[used])
    assert np.all((n["q"][:, 0, :] == 255)[~used]) and np.all((n["q"][:, 3, :] == 0)[~used])


def test_converter_rejects_malformed_input():
    flat = load_scene("cornell_org", "cpu")
    L = abi.lib()
    assert not L.imba_bvh8_build(7, flat.root, flat.nodes.ctypes.data, flat.tris.ctypes.data, flat.inst_nodes.ctypes.data)
    assert "layout" in abi.last_error()
    assert not L.imba_bvh8_build(0, flat.root, None, flat.tris.ctypes.data, flat.inst_nodes.ctypes.data)
    bad = flat.nodes.copy(); bad["children"][flat.root] = 0
    assert not L.imba_bvh8_build(0, flat.root, bad.ctypes.data, flat.tris.ctypes.data, flat.inst_nodes.ctypes.data)
