"""The reference's `.bvh` cache format (loaders/load_bvh.cpp): files written by the reference's own loader are read back and
must contain exactly the node / triangle arrays the flattened scenes hold; a new block type carries the 8-wide layout."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from imbatracer_b200 import abi
from tests import refapi
from tests.conftest import GOLDEN, load_scene


def read_block(path, btype, node_dt, prim_dt):
    L = abi.lib()
    f = L.imba_bvh_cache_open(path.encode())
    if not f:
        raise RuntimeError(abi.last_error())
    pn, pp, nn, np_ = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_uint32()
    rc = L.imba_bvh_cache_block(f, btype, pn, nn, pp, np_)
    if rc != 0:
        L.imba_bvh_cache_close(f); raise KeyError(abi.last_error())
    nodes = np.frombuffer((C.c_uint8 * (nn.value * node_dt.itemsize)).from_address(pn.value), node_dt).copy()
    prims = np.frombuffer((C.c_uint8 * (np_.value * prim_dt.itemsize)).from_address(pp.value), prim_dt).copy()
    L.imba_bvh_cache_close(f)
    return nodes, prims


@pytest.mark.parametrize("obj,scene", [("CornellBox-Original", "cornell_org"), ("CornellBox-Specular", "cornell_specular_front")])
def test_reads_caches_written_by_the_reference(obj, scene):
    path = os.path.join(GOLDEN, obj + ".bvh")
    vec4 = np.dtype(("<f4", 4))
    for btype, layout, node_dt in ((1, "gpu", abi.NODE2), (2, "cpu", abi.NODE4)):
        nodes, prims = read_block(path, btype, node_dt, vec4)
        flat = load_scene(scene, layout)
        # single mesh at offset 0: file-local indices are the scene's; the top-level nodes follow the mesh nodes
        assert nodes.tobytes() == flat.nodes[:len(nodes)].tobytes()
        assert len(nodes) == flat.root
        assert prims.tobytes() == flat.tris.reshape(-1, 4)[:len(prims)].tobytes()


def test_new_block_type_roundtrip_and_duplicate_refusal(tmp_path):
    src = os.path.join(GOLDEN, "CornellBox-Original.bvh")
    path = str(tmp_path / "c.bvh"); shutil.copy(src, path)
    L = abi.lib()
    b8 = refapi.Bvh8Host(load_scene("cornell_org", "gpu"))
    abi.check(L.imba_bvh_cache_append(path.encode(), 3, b8.nodes.ctypes.data, len(b8.nodes), b8.tris.ctypes.data, len(b8.tris)))
    nodes, tris = read_block(path, 3, abi.NODE8, abi.TRI48)
    assert nodes.tobytes() == b8.nodes.tobytes() and tris.tobytes() == b8.tris.tobytes()
    # the reference's blocks are still there and unchanged; a second block of the same type is refused (load_bvh.cpp:149-155)
    n1, p1 = read_block(path, 1, abi.NODE2, np.dtype(("<f4", 4)))
    n0, p0 = read_block(src, 1, abi.NODE2, np.dtype(("<f4", 4)))
    assert n1.tobytes() == n0.tobytes() and p1.tobytes() == p0.tobytes()
    assert L.imba_bvh_cache_append(path.encode(), 3, b8.nodes.ctypes.data, len(b8.nodes), b8.tris.ctypes.data, len(b8.tris)) != 0
    # a fresh file gets the magic; garbage is rejected
    fresh = str(tmp_path / "new.bvh")
    abi.check(L.imba_bvh_cache_append(fresh.encode(), 1, n0.ctypes.data, len(n0), p0.ctypes.data, len(p0)))
    assert open(fresh, "rb").read(4) == (0x313F1A57).to_bytes(4, "little")
    bad = str(tmp_path / "bad.bvh"); open(bad, "wb").write(b"not a cache")
    assert not L.imba_bvh_cache_open(bad.encode()) and "magic" in abi.last_error()


def test_append_never_destroys_an_existing_cache(tmp_path):
    """A cache whose LATER block is damaged still holds the good blocks the reference's loader wrote: appending decides "exists" from
    the magic alone, like store_accel_* (load_bvh.cpp:144-157), and appends -- it must not reopen the file for writing.  A file that
    is not a cache at all is refused, not overwritten; header counts larger than the file never size an allocation."""
    src = os.path.join(GOLDEN, "CornellBox-Original.bvh")
    good = open(src, "rb").read()
    L = abi.lib()
    b8 = refapi.Bvh8Host(load_scene("cornell_org", "gpu"))
    # good reference blocks followed by a block header that claims more payload than there is (a torn write)
    torn = str(tmp_path / "torn.bvh")
    open(torn, "wb").write(good + (1 << 20).to_bytes(8, "little") + (2).to_bytes(4, "little") + b"\x01\x02\x03")
    assert not L.imba_bvh_cache_open(torn.encode())                       # the strict reader refuses the file ...
    abi.check(L.imba_bvh_cache_append(torn.encode(), 3, b8.nodes.ctypes.data, len(b8.nodes), b8.tris.ctypes.data, len(b8.tris)))
    assert open(torn, "rb").read()[:len(good)] == good                   # ... but the append kept every byte that was there
    # not a cache: refused, untouched
    other = str(tmp_path / "notes.txt"); open(other, "wb").write(b"important, not a cache")
    assert L.imba_bvh_cache_append(other.encode(), 3, b8.nodes.ctypes.data, len(b8.nodes), b8.tris.ctypes.data, len(b8.tris)) != 0
    assert open(other, "rb").read() == b"important, not a cache"
    # absurd counts in a header: error, not an allocation of 4 G nodes
    huge = str(tmp_path / "huge.bvh")
    payload = 8 + 56 * 0xFFFFFFF0 + 16 * 0xFFFFFFF0
    open(huge, "wb").write((0x313F1A57).to_bytes(4, "little") + (payload + 4).to_bytes(8, "little") + (1).to_bytes(4, "little") +
                           (0xFFFFFFF0).to_bytes(4, "little") * 2)
    assert not L.imba_bvh_cache_open(huge.encode()) and "larger than the file" in abi.last_error()
