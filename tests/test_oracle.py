"""CPU tests that PIN the oracle (oracle/traversal_oracle.cpp): against committed goldens, against brute force,
between the two reference layouts, against the reference's own RNG, and (when the reference build is present)
against the reference's converged golden images."""
import hashlib
import os

import numpy as np
import pytest

from imbatracer_b200 import abi
from tests import refapi
from tests.conftest import GOLDEN, SCENE_NAMES, batch_for, load_hits, load_scene


@pytest.mark.parametrize("name", SCENE_NAMES)
@pytest.mark.parametrize("layout", ["cpu", "gpu"])
def test_oracle_matches_committed_hits(name, layout):
    flat = load_scene(name, layout)
    rays = batch_for(name, flat)
    g = load_hits(name)
    hits = refapi.trace_oracle(flat, rays)
    assert hits[:4096].tobytes() == g[f"closest_{layout}_head"].tobytes()
    assert hashlib.sha256(hits.tobytes()).digest() == g[f"closest_{layout}_sha"].tobytes()
    occ = refapi.trace_oracle(flat, rays, any_hit=True)
    assert np.array_equal(np.packbits(occ["tri_id"] >= 0), g[f"occluded_{layout}"])


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_oracle_layouts_agree(name):
    """MBVH4 and BVH2 trees of the same scene differ in shape; the tie rule makes the result shape-independent."""
    a, b = load_scene(name, "cpu"), load_scene(name, "gpu")
    rays = batch_for(name, a)
    ha, hb = refapi.trace_oracle(a, rays), refapi.trace_oracle(b, rays)
    assert ha.tobytes() == hb.tobytes()


def brute_force(flat, rays):
    """Numpy restatement of the triangle test over ALL triangles (single-instance scenes), float32 throughout."""
    tris = flat.tris.reshape(-1, 4)
    assert flat.accel_layout == 1
    # collect triangles from the BVH2 leaf array: triples (v0, v1[w=id], v2); duplicates are harmless
    n = len(tris) // 3
    v0, v1, v2 = tris[0:3 * n:3, :3], tris[1:3 * n:3, :3], tris[2:3 * n:3, :3]
    ids = tris[1:3 * n:3, 3].copy().view(np.int32)
    f = np.float32
    e1, e2 = (v0 - v1).astype(f), (v2 - v0).astype(f)
    nn = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                   e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], 1).astype(f)
    out = np.zeros(len(rays), abi.HIT)
    for i, r in enumerate(rays):
        o, d = r["org"].astype(np.float64), r["dir"].astype(np.float64)
        c = v0.astype(np.float64) - o
        rr = np.cross(d, c)
        det = nn.astype(np.float64) @ d
        sgn = np.where(np.signbit(det), -1.0, 1.0)
        U = (rr * e2).sum(1) * sgn; V = (rr * e1).sum(1) * sgn; T = (nn * c).sum(1) * sgn
        adet = np.abs(det)
        ok = (U >= 0) & (V >= 0) & (U + V <= adet) & (det != 0) & (T >= adet * r["tmin"]) & (T <= adet * np.float64(r["tmax"]))
        if not ok.any():
            out[i] = (-1, -1, r["tmax"], 0)
            continue
        t = np.where(ok, T / np.where(adet == 0, 1, adet), np.inf)
        k = int(np.argmin(t))
        out[i] = (ids[k], 0, t[k], U[k] / adet[k])
    return out


def test_oracle_vs_brute_force():
    """Away from exact ties and silhouettes the BVH-accelerated oracle must find the nearest of ALL triangles."""
    flat = load_scene("cornell_specular_front", "gpu")
    rays = batch_for("cornell_specular_front", flat)[::37][:600]
    h = refapi.trace_oracle(flat, rays)
    b = brute_force(flat, rays)
    hit_o, hit_b = h["tri_id"] >= 0, b["tri_id"] >= 0
    # float64 brute force vs float32 oracle: allow disagreement only where the hit is a near-tie / silhouette case
    agree = (hit_o == hit_b) & (~hit_o | (np.abs(h["tmax"] - b["tmax"]) <= 1e-5 * np.maximum(1.0, np.abs(b["tmax"]))))
    assert agree.mean() > 0.995, agree.mean()
    same_tri = (h["tri_id"] == b["tri_id"]) | ~hit_o | ~hit_b
    assert same_tri.mean() > 0.98


def mwc64x(state, n):
    """render/random.h:15-47 restated: x*4294883355 + c, output x^c, float = u32 / 0xFFFFFFFF."""
    out = []
    for _ in range(n):
        c, x = state >> 32, state & 0xFFFFFFFF
        state = (x * 4294883355 + c) & 0xFFFFFFFFFFFFFFFF
        out.append(np.float32(np.float32(np.uint32(x ^ c)) / np.float32(0xFFFFFFFF)))
    return np.array(out, np.float32), state


def test_rng_known_answers():
    k = np.load(os.path.join(GOLDEN, "rng_kat.npz"))
    # published probe of the survey: RNG(12345) produces these three values first
    assert set(np.round(k["floats_12345"][:3].astype(np.float64), 6)) == {0.282635, 0.758732, 0.000003}
    assert list(k["struct_sizes"]) == [32, 16, 16, 16, 40, 48, 16, 88, 120, 44, 112, 56, 64]


@pytest.mark.skipif(not refapi.have_ref() or not os.path.isdir(refapi.REF_SCENES), reason="reference build not present")
@pytest.mark.parametrize("scene,algo,radius,iters,tol", [("cornell_org", 0, 2.0, 48, 0.04), ("cornell_indirect", 1, 2.0, 48, 0.03), ("cornell_specular", 2, 0.003, 64, 0.08),
                                                        ("cornell_specular_front", 2, 0.003, 64, 0.07), ("cornell_water", 2, 0.003, 64, 0.11)])
def test_reference_pipeline_with_oracle_reproduces_golden_image(scene, algo, radius, iters, tol):
    """The reference's own integrators running on the oracle's traversal must converge to all five usable golden PNGs of the
    reference with the integrator and radius test/render_refs.sh:10-23 used for each (PT, BPT, VCM -r 0.003), at 128x128,
    gamma-0.5 RMSE as test/run_bench.py:113-129 measures it.  Tolerances = what 48 / 64 samples per pixel leave at this size
    (0.025 / 0.017 / 0.061 / 0.051 / 0.090 measured); the three VCM scenes keep a bias of ~0.05 at 128^2 because the golden PNGs
    clamp their caustics per 1024^2 pixel (the device renders are gated at 1024^2 and 0.03, tests/test_goldens_gpu.py)."""
    rs = refapi.RefScene(refapi.scene_path(scene), False)
    img, st = rs.render(algo, 128, 128, iterations=iters, radius_factor=radius, tile=64)
    rs.close()
    got = np.clip(img / float(iters), 0, None) ** 0.5
    want = np.load(os.path.join(GOLDEN, f"ref_{scene}_128.npy")) ** 0.5
    rmse = float(np.sqrt(np.mean((np.clip(got, 0, 1) - np.clip(want, 0, 1)) ** 2)))
    assert rmse < tol, rmse


def test_coplanar_rays_are_the_known_hit_or_miss_class():
    """DESIGN.md section 1, second degenerate class: two shadow rays of the bench's incoherent batch that lie in the plane of
    the emitter triangle they 'hit'.  det = n.d is rounding noise, the computed hit point lies outside the triangle's bounds;
    the exact boxes of the reference layout cull the leaf (oracle: miss), the quantised 8-wide boxes do not (walker and CUDA
    path on the converted tree: hit)."""
    from imbatracer_b200 import procedural
    flat = procedural.still_life_proxy(1.0)
    rays = np.zeros(2, abi.RAY)
    rays[0] = ((1.8839427, 7.1488204, 4.8558984), 0.001, (-0.02364074, -0.045666218, 0.2283302), 0.999)
    rays[1] = ((3.3762798, 6.1687307, -0.2962015), 0.001, (-0.09021112, -0.20622504, -0.05316282), 0.999)
    assert (refapi.trace_oracle(flat, rays)["tri_id"] == -1).all() and (refapi.trace_oracle(flat, rays, any_hit=True)["tri_id"] == -1).all()
    w = refapi.Bvh8Host(flat).walk(rays)
    assert w["tri_id"].tolist() == [1516449, 1516456]
    idx = flat.indices.reshape(-1, 4); v = flat.vertices.reshape(-1, 4)[:, :3].astype(np.float64)
    for r, h in zip(rays, w):
        p0, p1, p2 = v[idx[h["tri_id"], :3]]
        n = np.cross(p0 - p1, p2 - p0); d = r["dir"].astype(np.float64)
        assert abs(n @ d) < 1e-6 * np.linalg.norm(n) * np.linalg.norm(d)          # the ray lies in the triangle's plane
        p = r["org"].astype(np.float64) + float(h["tmax"]) * d
        lo, hi = np.minimum(np.minimum(p0, p1), p2), np.maximum(np.maximum(p0, p1), p2)
        assert ((p < lo - 1e-3) | (p > hi + 1e-3)).any()                          # the reported point is not on the triangle


def test_edge_of_box_ray_is_the_known_hit_or_miss_class():
    """DESIGN.md section 1, stress run: a ray that crosses the plane of a wall triangle 5e-5 ABOVE its top edge.  The inside
    test's rounding still accepts the point; the exact leaf box of the reference layout ends at the edge (oracle: miss), the
    8-bit boxes of the converted tree are a rounding wider (walker and CUDA path: hit).  Found by tools/gpu_parity_stress.py."""
    from imbatracer_b200 import procedural
    flat = procedural.still_life_proxy(1.0)
    rays = np.zeros(1, abi.RAY)
    rays[0] = ((-2.0397363, 8.315759, 3.6616228), 0.0, (-0.4574975, -0.073693514, 0.886152), np.finfo(np.float32).max)
    assert refapi.trace_oracle(flat, rays)["tri_id"][0] == -1
    w = refapi.Bvh8Host(flat).walk(rays)
    assert w["tri_id"][0] == 0 and abs(float(w["tmax"][0]) - 4.284753) < 1e-5
    idx = flat.indices.reshape(-1, 4); v = flat.vertices.reshape(-1, 4)[:, :3].astype(np.float64)
    tri = v[idx[0, :3]]
    p = rays[0]["org"].astype(np.float64) + float(w["tmax"][0]) * rays[0]["dir"].astype(np.float64)
    over = p[1] - tri[:, 1].max()
    assert 0 < over < 1e-4           # outside the triangle's own bounds, by a rounding error of the ray arithmetic

