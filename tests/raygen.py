"""TEST INFRASTRUCTURE: seeded ray batches shared by the fixture generator and the parity tests."""
import numpy as np

from imbatracer_b200 import abi


def random_rays(scene, n, seed, tmax_mix=True):
    """Incoherent rays: origins in a box around the scene's bounding sphere, uniform directions."""
    rng = np.random.default_rng(seed)
    r = np.zeros(n, abi.RAY)
    c, rad = scene.bsphere[:3].astype(np.float32), np.float32(scene.bsphere[3])
    r["org"] = (c + (rng.random((n, 3), np.float32) - np.float32(0.5)) * rad * np.float32(1.2)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    r["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    r["tmin"] = np.where(rng.random(n) < 0.2, rng.random(n) * 0.1 * rad, 0.0).astype(np.float32)
    if tmax_mix:
        r["tmax"] = np.where(rng.random(n) < 0.3, rng.random(n) * rad, np.finfo(np.float32).max).astype(np.float32)
    else:
        r["tmax"] = np.finfo(np.float32).max
    return r


def edge_rays(scene, vertex_rays=True):
    """Padding rays (all zero, ray_queue.h:74,250), axis-parallel rays with exact zeros of both signs, tmax = 0,
    rays starting on vertices, unnormalised directions."""
    c, rad = scene.bsphere[:3].astype(np.float32), np.float32(scene.bsphere[3])
    out = [np.zeros(64, abi.RAY)]
    ax = np.zeros(12 * 16, abi.RAY)
    k = 0
    for axis in range(3):
        for sgn in (1.0, -1.0):
            for neg_zero in (False, True):
                for j in range(8):
                    o = c.copy(); o[(axis + 1) % 3] += (j - 3.5) * 0.11 * rad; o[(axis + 2) % 3] += (j % 3 - 1) * 0.17 * rad
                    o[axis] -= sgn * rad * 1.5
                    d = np.zeros(3, np.float32); d[axis] = sgn
                    if neg_zero:
                        d[(axis + 1) % 3] = -0.0; d[(axis + 2) % 3] = -0.0
                    ax[k]["org"] = o; ax[k]["dir"] = d; ax[k]["tmax"] = np.finfo(np.float32).max; k += 1
    out.append(ax[:k])
    v = scene.vertices.reshape(-1, 4)[:, :3]
    nv = min(len(v), 512) if vertex_rays else 0
    rng = np.random.default_rng(7)
    vr = np.zeros(nv, abi.RAY)
    vr["org"] = v[:nv]
    d = rng.normal(size=(nv, 3)).astype(np.float32) * np.float32(3.7)  # un-normalised on purpose
    vr["dir"] = d; vr["tmax"] = np.finfo(np.float32).max
    out.append(vr)
    z = random_rays(scene, 128, 99)
    z["tmax"] = 0.0
    out.append(z)
    return np.concatenate(out)


def test_batch(scene, camera_rays):
    return np.concatenate([camera_rays, random_rays(scene, 16384, 1234), edge_rays(scene)])
