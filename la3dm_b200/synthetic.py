"""Seeded synthetic range scans for the headline benchmark.

"64 k-point scans into a 50 m / 0.1 m map": an axis-aligned room of edge `extent` (height extent/8) with 200 random
spheres as unstructured clutter; `n_points` isotropic rays from a sensor near the centre, first hit + N(0, 0.01^2) noise,
float32.  Successive scans move the sensor on a seeded random walk (~1 m steps) so that scans overlap (accumulation and
pruning are exercised).  Pure numpy so that bench.py can regenerate the identical workload on the GPU box.
"""
import numpy as np


class Scene:
    def __init__(self, extent=50.0, n_spheres=200, seed=1):
        rng = np.random.default_rng(seed)
        L = float(extent)
        self.lo = np.array([-L / 2, -L / 2, 0.0])
        self.hi = np.array([L / 2, L / 2, L / 8])
        self.centers = np.c_[rng.uniform(-L / 2, L / 2, (n_spheres, 2)), rng.uniform(0, L / 8, n_spheres)]
        self.radii = rng.uniform(0.3, 1.5, n_spheres)
        self.extent = L

    def scan(self, origin, n_points, rng):
        o = np.asarray(origin, np.float64)
        d = rng.normal(size=(n_points, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (self.lo - o) / d
            t2 = (self.hi - o) / d
        t = np.where(d > 0, t2, t1).min(1)
        for c, r in zip(self.centers, self.radii):
            oc = o - c
            b = d @ oc
            cc = oc @ oc - r * r
            disc = b * b - cc
            th = np.where(disc > 0, -b - np.sqrt(np.maximum(disc, 0)), np.inf)
            th = np.where(th > 0.3, th, np.inf)
            t = np.minimum(t, th)
        p = o + d * t[:, None] + rng.normal(scale=0.01, size=(n_points, 3))
        return p.astype(np.float32)


def make_sequence(n_scans=12, n_points=65536, extent=50.0, seed=1):
    """-> (points [n_scans, n_points, 3] float32, origins [n_scans, 3] float32)"""
    scene = Scene(extent, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    o = np.array([0.05, 0.03, 1.07])
    pts, org = [], []
    for _ in range(n_scans):
        org.append(o.astype(np.float32))
        pts.append(scene.scan(org[-1].astype(np.float64), n_points, rng))
        step = rng.normal(size=3) * np.array([0.7, 0.7, 0.05])
        o = o + step
        o[:2] = np.clip(o[:2], -extent / 2 + 2, extent / 2 - 2)
        o[2] = float(np.clip(o[2], 0.5, extent / 8 - 0.5))
    return np.stack(pts), np.stack(org)
