"""Generate tests/golden/regions.npz from the UNMODIFIED reference's geometry objects:

    python tests/golden/make_region_golden.py

For each problem class the full training geometry (``pde.geomtime`` or ``pde.geom``, what ``create_model`` hands to
``dde.data.PDE / TimePDE``, src/pde/baseclass.py:120-140) is
  * turned into a sampling region by ``pinnacle_b200.resample.geometry_region`` (stored: lo, hi, holes), and
  * probed with the reference's own code: ``geom.inside(p)`` on 4096 box-uniform probe points (stored with the verdicts), and
    256 points of ``geom.random_points(256, random="pseudo")`` (stored; the reference's own draw).
The tests check the region's point-in-hole rule against the stored verdicts (oracle on CPU, kernel on the GPU) without the
reference tree.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402
from pinnacle_b200.resample import geometry_region  # noqa: E402

CASES = ["Burgers1D", "Poisson2D_Classic", "Heat2D_ComplexGeometry", "Heat2D_Multiscale", "NS2D_Classic", "NS2D_LidDriven", "NS2D_BackStep",
         "Poisson3D_ComplexGeometry", "KuramotoSivashinskyEquation", "Wave1D", "PoissonBoltzmann2D", "GrayScottEquation", "HeatND", "PoissonND"]


def main():
    refshim.load()
    out = {}
    meta = {}
    for name in CASES:
        pde = refshim.construct(name)
        geom = getattr(pde, "geomtime", None) or pde.geom
        reg = geometry_region(geom)
        rng = np.random.default_rng(sum(map(ord, name)))
        lo, hi = np.asarray(reg.lo), np.asarray(reg.hi)
        probe = (lo + rng.random((4096, reg.d)) * (hi - lo)).astype(np.float32)
        # 24 probes per hole are pulled towards the hole boundaries, where the verdict is decided
        for i, (s, a, b) in enumerate(reg.holes):
            k = len(a)
            sel = slice(i * 24, i * 24 + 24)
            ang = rng.normal(size=(24, k))
            ang /= np.linalg.norm(ang, axis=1, keepdims=True)
            if s == 0:
                probe[sel, :k] = (np.asarray(a) + ang * b[0] * (1 + rng.uniform(-0.02, 0.02, (24, 1)))).astype(np.float32)
        if type(geom).__name__ == "GeometryXTime":       # no inside() of its own: the spatial verdict (t is in range by construction)
            inside = np.asarray(geom.geometry.inside(probe[:, :-1])).astype(bool)
        else:
            inside = np.asarray(geom.inside(probe)).astype(bool)
        if reg.keep is not None:                          # pull 64 probes onto the sphere that bounds the domain
            k = len(reg.keep[0])
            ang = rng.normal(size=(64, k))
            ang /= np.linalg.norm(ang, axis=1, keepdims=True)
            probe[-64:, :k] = (np.asarray(reg.keep[0]) + ang * reg.keep[1] * (1 + rng.uniform(-0.02, 0.02, (64, 1)))).astype(np.float32)
            inside = np.asarray(geom.geometry.inside(probe[:, :-1]) if type(geom).__name__ == "GeometryXTime" else geom.inside(probe)).astype(bool)
        np.random.seed(sum(map(ord, name)))
        own = np.asarray(geom.random_points(256, random="pseudo"), dtype=np.float32)
        out[name + "_probe"] = probe
        out[name + "_inside"] = inside
        out[name + "_refdraw"] = own
        meta[name] = dict(lo=list(reg.lo), hi=list(reg.hi), holes=[[s, list(a), list(b)] for s, a, b in reg.holes], geometry=type(geom).__name__,
                          keep=None if reg.keep is None else [list(reg.keep[0]), reg.keep[1]])
        print(f"{name:32s} d={reg.d} holes={len(reg.holes):2d} inside fraction {inside.mean():.3f}")
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "regions.npz"), **out)


if __name__ == "__main__":
    main()
