"""TEST INFRASTRUCTURE -- CPU restatement of the device point sampler (``pinn_sample_points``, include/pinn_b200.h).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this module; the product path
(``pinnacle_b200/resample.py``) never does.

What it follows.  The reference redraws residual points with ``geom.random_points(n, random="pseudo")``
(deepxde/data/pde.py:232-234 under ``PDE.resample_train_points`` :186-193): ``np.random.rand(n, dim)`` scaled into the
bounding box (deepxde/geometry/geometry_nd.py Hypercube.random_points; geometry_2d.py Rectangle inherits it), times a uniform
``t`` for ``GeometryXTime`` (timedomain.py:86-98 with "pseudo"), with rejection of the points inside a subtracted shape
(deepxde/geometry/csg.py CSGDifference.random_points: draw from geom1, keep ``~geom2.inside``, repeat until n points); a
``Hypersphere`` domain (HeatND, src/pde/heat.py:257) is sampled by the reference as Gaussian direction x radius**(1/d)
(geometry_nd.py:138-150), here by rejection from its bounding cube -- the same uniform law on the ball.  The
distribution is the same; the random stream is not numpy's MT19937 but a counter-based one (Philox4x32-10, Salmon et al.,
"Parallel random numbers: as easy as 1, 2, 3", SC'11 -- the generator is pinned below by the known-answer vectors of that
paper's Random123 distribution), so that every row is a pure function of (key, draw, row index) and the device can produce
the rows independently.  Parity with the reference is therefore distributional (tests/test_resample.py: the reference
geometry's own ``inside()`` accepts every point, moments of the box-uniform law); parity device <-> this file is bit-exact.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
MAX_ATTEMPTS = 256
HOLE_BALL, HOLE_BOX = 0, 1


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Ten rounds of Philox-4x32 on arrays of counters (uint32 each) with the scalar key (k0, k1)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for r in range(10):
        p0, p1 = M0 * c0, M1 * c2                      # 32 x 32 -> 64 bit products
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        ka, kb = np.uint64((k0 + r * W0) & 0xFFFFFFFF), np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        c0, c1, c2, c3 = hi1 ^ c1 ^ ka, lo1, hi0 ^ c3 ^ kb, lo0
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _draw(rows: np.ndarray, attempt: np.ndarray, lo: np.ndarray, hi: np.ndarray, key: int, draw: int) -> np.ndarray:
    d = lo.shape[0]
    out = np.empty((rows.shape[0], d), dtype=np.float32)
    span = (hi - lo).astype(np.float32)                # float32 subtraction, as the device does
    for b in range((d + 3) // 4):
        w = philox4x32_10(rows & MASK, rows >> np.uint64(32), attempt, np.uint64(b | (draw << 8)), key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF)
        for j in range(4):
            c = 4 * b + j
            if c < d:
                u = (w[j] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
                out[:, c] = (u * span[c]).astype(np.float32) + lo[c]
    return out


def in_holes(p: np.ndarray, holes) -> np.ndarray:
    bad = np.zeros(p.shape[0], dtype=bool)
    for shape, a, b in holes:
        a = np.asarray(a, dtype=np.float32)
        b = np.asarray(b, dtype=np.float32)
        if shape == HOLE_BALL:
            s = np.zeros(p.shape[0], dtype=np.float32)
            for j in range(a.shape[0]):
                dj = p[:, j] - a[j]
                s = s + dj * dj                        # float32 products and sums in the device's order
            bad |= s <= b[0] * b[0]
        else:
            inside = np.ones(p.shape[0], dtype=bool)
            for j in range(a.shape[0]):
                inside &= (p[:, j] >= a[j]) & (p[:, j] <= b[j])
            bad |= inside
    return bad


def outside_ball(p: np.ndarray, keep) -> np.ndarray:
    """keep = (centre over the first k coordinates, radius) or None: True where the point lies outside the closed ball
    (deepxde/geometry/geometry_nd.py Hypersphere.inside: ``norm(x - center) <= radius``)."""
    if keep is None:
        return np.zeros(p.shape[0], dtype=bool)
    c, r = np.asarray(keep[0], dtype=np.float32), np.float32(keep[1])
    s = np.zeros(p.shape[0], dtype=np.float32)
    for j in range(c.shape[0]):
        dj = p[:, j] - c[j]
        s = s + dj * dj
    return s > r * r


def sample_points(n_rows: int, lo, hi, holes=(), key: int = 0, draw: int = 0, first_row: int = 0, keep=None) -> np.ndarray:
    """Rows first_row .. first_row + n_rows of the stream (key, draw): float32 [n_rows, d].  ``holes``: sequence of
    (shape, a, b) as in ``pinn_hole``; ``keep``: (centre, radius) of the ball the points must lie in, or None."""
    lo = np.asarray(lo, dtype=np.float32)
    hi = np.asarray(hi, dtype=np.float32)
    rows = np.arange(n_rows, dtype=np.uint64) + np.uint64(first_row)
    out = np.empty((n_rows, lo.shape[0]), dtype=np.float32)
    todo = np.arange(n_rows)
    for attempt in range(MAX_ATTEMPTS):
        if todo.size == 0:
            break
        p = _draw(rows[todo], np.uint64(attempt), lo, hi, key, draw)
        out[todo] = p                                   # the last draw stays if every attempt is rejected
        todo = todo[outside_ball(p, keep) | in_holes(p, holes)]
    return out
