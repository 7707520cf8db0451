"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the header declares,
and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from imbatracer_b200 import abi
from tests.conftest import ROOT, load_scene


def declared_functions():
    src = open(os.path.join(ROOT, "include", "imba_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\(", src)
    return sorted({n for n in names if n.startswith("imba_") or n.endswith("_masked_instanced")})


def test_library_exports_every_declared_symbol():
    L = C.CDLL(abi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    flat = load_scene("cornell_org", "cpu")
    d = flat.desc()
    h = abi.lib().imba_scene_create(C.byref(d))
    assert not h
    assert "no CUDA device" in abi.last_error()


def test_oracle_is_not_linked_into_the_product():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out and "bvh8_walk" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "imbatracer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")) :
                text = open(os.path.join(root, f)).read()
                assert "refapi" not in text and "liboracle" not in text, f


def test_primary_ray_seeding_matches_reference_known_answers():
    """ray_gen.h:57-72 (Bernstein hash in int arithmetic, sign-extended seed, discard count) -- host code, no GPU."""
    import numpy as np
    from tests.conftest import GOLDEN
    k = np.load(os.path.join(GOLDEN, "rng_kat.npz"))
    out = np.zeros(4096, np.uint64)
    abi.check(abi.lib().imba_seed_states(0x1234ABCD, 0, 4096, 4, out.ctypes.data))
    assert np.array_equal(out, k["seed_states"])


def test_cpp_host_mirror_compiles_links_and_refuses_without_gpu(tmp_path):
    """include/imba_b200.hpp (RayQueue / Integrator mirror of ray_queue.h, integrator.h) builds with the host compiler and
    links against the product library; on a box without a GPU the library refuses instead of falling back."""
    import subprocess
    exe = str(tmp_path / "host_mirror_test")
    lib_dir = os.path.dirname(abi.LIB_PATH)
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                    "-o", exe, "-L", lib_dir, "-limba_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode in (0, 2), (r.returncode, r.stdout)


def test_traversal_kernel_resource_budget():
    """The persistent traversal kernel is tuned for 56 registers (9 resident blocks of 128 threads per SM) and no spills;
    ptxas writes its resource usage to the build log.  A change that silently pushes it to 90+ registers halves occupancy."""
    import os
    import re
    log = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "imbatracer_b200", "csrc", "build", "traverse.ptxas.log")
    if not os.path.exists(log):
        import pytest
        pytest.skip("library was not built in this tree")
    text = open(log).read()
    regs = [int(x) for x in re.findall(r"Used (\d+) registers", text)]
    spills = [int(x) for x in re.findall(r"(\d+) bytes spill stores", text)]
    assert regs and max(regs) <= 64, regs
    assert max(spills) == 0, spills
    # the flat closest-hit / any-hit kernels (no counters) are tuned for 48 registers = 10 resident blocks
    flat = re.findall(r"Compiling entry function '\S*trace_flat_kernelILi([01])ELb([01])ELb0E\S*'.*?Used (\d+) registers", text, re.S)
    assert len(flat) == 4 and all(int(r) <= 48 for _, _, r in flat), flat
