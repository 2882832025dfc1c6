"""Residual points redrawn on the device (``PDEPointResampler`` on device).

``dde.callbacks.PDEPointResampler`` (deepxde/callbacks.py:540-575) redraws the training points every ``period`` iterations
on the host -- ``PDE.resample_train_points`` (deepxde/data/pde.py:186-193) -> ``geom.random_points(num_domain, "pseudo")`` --
and the next train step uploads the new array.  ``DevicePointResampler`` keeps the schedule and the distribution (uniform
over the problem's geometry: bounding box minus the subtracted disks / spheres / rectangles, times the time interval) but
draws the ``num_domain`` rows with one kernel launch (``pinn_sample_points``) straight into the staged device array of the
fused op; the per-point source / coefficient columns are re-evaluated on the device as well.  Nothing crosses PCIe until
somebody needs the host copy (``sync_host()``, called at the end of ``train()``).

Differences from the reference's callback, all deliberate:

* the random stream is Philox4x32-10 keyed by (seed, number of the redraw, row), not numpy's global MT19937 state: a run is
  reproducible from ``seed`` alone and does not depend on how many ranks share the rows;
* only the ``num_domain`` interior rows are redrawn; the boundary / initial rows that deepxde also keeps in the residual
  set (``train_points()``, data/pde.py:236-243, :312-319) stay (``bc_points`` resampling is not offered: it would change
  the boundary terms' targets, which are evaluated once per point set);
* geometries outside the (box or ball)-minus-holes family (polygons, unions as domains) are refused at ``on_train_begin``;
  a ``Hypersphere`` domain (HeatND) is drawn by rejection from its bounding cube (the reference: Gaussian direction times
  radius**(1/d), deepxde/geometry/geometry_nd.py:138-150 -- the same uniform law).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import capi
from .kinds import HOLE_BALL, HOLE_BOX, PINN_MAX_DIM, PINN_MAX_HOLES, CRegion

Hole = Tuple[int, Tuple[float, ...], Tuple[float, ...]]      # (shape, a, b) as in struct pinn_hole


@dataclass
class Region:
    """Axis-aligned box [lo, hi), optionally intersected with one closed ball over the leading coordinates (``keep`` =
    (centre, radius): a ``Hypersphere`` domain inside its bounding cube), minus closed holes (``struct pinn_region``)."""
    lo: Sequence[float]
    hi: Sequence[float]
    holes: List[Hole] = field(default_factory=list)
    keep: Optional[Tuple[Tuple[float, ...], float]] = None

    def __post_init__(self):
        self.lo = tuple(float(v) for v in self.lo)
        self.hi = tuple(float(v) for v in self.hi)
        if len(self.lo) != len(self.hi) or not 1 <= len(self.lo) <= PINN_MAX_DIM:
            raise ValueError(f"region: need 1..{PINN_MAX_DIM} coordinates, lo and hi of the same length")
        if any(not h >= l for l, h in zip(self.lo, self.hi)):
            raise ValueError("region: hi < lo")
        if len(self.holes) > PINN_MAX_HOLES:
            raise ValueError(f"region: more than {PINN_MAX_HOLES} holes")
        if self.keep is not None:
            self.keep = (tuple(float(v) for v in self.keep[0]), float(self.keep[1]))
            if not 1 <= len(self.keep[0]) <= self.d or not self.keep[1] > 0:
                raise ValueError("region: the kept ball spans 1..d coordinates and has a positive radius")
        self.holes = [(int(s), tuple(float(v) for v in a), tuple(float(v) for v in b)) for s, a, b in self.holes]
        for s, a, b in self.holes:
            if s not in (HOLE_BALL, HOLE_BOX) or not 1 <= len(a) <= min(3, self.d) or (s == HOLE_BOX and len(b) != len(a)) or len(b) < 1:
                raise ValueError(f"region: bad hole {(s, a, b)}")

    @property
    def d(self) -> int:
        return len(self.lo)

    def to_c(self) -> CRegion:
        r = CRegion()
        r.d, r.n_holes = self.d, len(self.holes)
        for j in range(self.d):
            r.lo[j], r.hi[j] = self.lo[j], self.hi[j]
        if self.keep is not None:
            r.keep_dims, r.keep_radius = len(self.keep[0]), self.keep[1]
            for j, v in enumerate(self.keep[0]):
                r.keep_center[j] = v
        for i, (s, a, b) in enumerate(self.holes):
            r.holes[i].shape, r.holes[i].n_dims = s, len(a)
            for j, v in enumerate(a):
                r.holes[i].a[j] = v
            for j, v in enumerate(b[:3]):
                r.holes[i].b[j] = v
        return r

    def volume_fraction_bound(self) -> float:
        """Upper bound of the fraction of the box that is rejected (outside the kept ball + the holes' volumes inside the
        spatial box, ignoring overlaps and clipping) -- a redraw takes ``attempts ~ 1 / (1 - fraction)`` Philox blocks per row."""
        tot = 0.0
        if self.keep is not None:
            import math
            k, r = len(self.keep[0]), self.keep[1]
            ball = math.pi ** (k / 2) / math.gamma(k / 2 + 1) * r ** k
            box = float(np.prod([self.hi[j] - self.lo[j] for j in range(k)]))
            tot += max(1.0 - ball / box, 0.0) if box > 0 else 0.0
        for s, a, b in self.holes:
            k = len(a)
            box = float(np.prod([self.hi[j] - self.lo[j] for j in range(k)]))
            if s == HOLE_BALL:
                r = b[0]
                v = {1: 2 * r, 2: np.pi * r * r, 3: 4.0 / 3.0 * np.pi * r ** 3}[k]
            else:
                v = float(np.prod([min(b[j], self.hi[j]) - max(a[j], self.lo[j]) for j in range(k)]))
            tot += max(v, 0.0) / box if box > 0 else 0.0
        return min(tot, 1.0)


def _cls(g) -> str:
    return type(g).__name__


def _holes_of(g) -> List[Hole]:
    """The closed set a subtracted deepxde geometry removes (csg.py CSGDifference.inside: ``geom1.inside & ~geom2.inside``)."""
    n = _cls(g)
    if n in ("Disk", "Sphere", "Hypersphere"):
        c = np.asarray(g.center, dtype=np.float64).reshape(-1)
        if c.shape[0] > 3:
            raise capi.FusedOpError(f"device resampler: a {c.shape[0]}-dimensional ball cannot be a hole (at most 3 coordinates)")
        return [(HOLE_BALL, tuple(c), (float(g.radius),))]
    if n == "Interval":
        return [(HOLE_BOX, (float(g.l),), (float(g.r),))]
    if n in ("Rectangle", "Cuboid", "Hypercube"):
        lo, hi = np.asarray(g.xmin, dtype=np.float64).reshape(-1), np.asarray(g.xmax, dtype=np.float64).reshape(-1)
        if lo.shape[0] > 3:
            raise capi.FusedOpError("device resampler: a box hole has at most 3 coordinates")
        return [(HOLE_BOX, tuple(lo), tuple(hi))]
    if n == "CSGUnion":
        return _holes_of(g.geom1) + _holes_of(g.geom2)
    raise capi.FusedOpError(f"device resampler: cannot subtract a {n}")


def geometry_region(geom) -> Region:
    """The sampling region of a deepxde geometry (duck-typed on the class names of deepxde/geometry and
    src/utils/geom.py, so that this module does not import deepxde)."""
    n = _cls(geom)
    if n == "GeometryXTime":
        r = geometry_region(geom.geometry)
        return Region(r.lo + (float(geom.timedomain.t0),), r.hi + (float(geom.timedomain.t1),), r.holes, r.keep)
    if n in ("Disk", "Sphere", "Hypersphere"):          # a ball as the DOMAIN (HeatND, src/pde/heat.py:257): its bounding cube, outside rejected
        c = np.asarray(geom.center, dtype=np.float64).reshape(-1)
        rad = float(geom.radius)
        return Region(c - rad, c + rad, [], (tuple(c), rad))
    if n in ("Interval", "TimeDomain"):
        return Region((float(geom.l),), (float(geom.r),))
    if n in ("Rectangle", "Cuboid", "Hypercube"):
        return Region(np.asarray(geom.xmin).reshape(-1), np.asarray(geom.xmax).reshape(-1))
    if n == "CSGDifference":
        r = geometry_region(geom.geom1)
        return Region(r.lo, r.hi, r.holes + _holes_of(geom.geom2), r.keep)
    if n == "CSGMultiDifference":                       # src/utils/geom.py:7-43
        r = geometry_region(geom.geom1)
        holes = list(r.holes)
        for g2 in geom.geom2_list:
            holes += _holes_of(g2)
        return Region(r.lo, r.hi, holes, r.keep)
    raise capi.FusedOpError(f"device resampler: geometry {n} is not a box or ball minus disks / spheres / boxes (times a time interval); "
                            "use dde.callbacks.PDEPointResampler for it")


def sample_region(region: Region, n_rows: int, key: int = 0, draw: int = 0, first_row: int = 0, device=None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``n_rows`` uniform points of ``region`` as a device tensor [n_rows, d] (rows first_row.. of the stream (key, draw))."""
    if out is None:
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        out = torch.empty((int(n_rows), region.d), dtype=torch.float32, device=dev)
    if out.shape[0] != n_rows:
        raise ValueError("sample_region: out has the wrong number of rows")
    if n_rows:
        capi.sample_points(out, region.to_c(), key, draw, first_row)
    return out


class DevicePointResampler:
    """Drop-in for ``dde.callbacks.PDEPointResampler(period, pde_points=True, bc_points=False)`` on a model patched by
    ``enable_fused``.  Duck-typed deepxde callback (the hooks ``CallbackList`` calls)."""

    def __init__(self, period: int = 100, seed: int = 0, region: Optional[Region] = None):
        self.period = int(period)
        self.seed = int(seed)
        self.region = region
        self.model = None
        self.epochs_since_last_resample = 0
        self.draws = 0                  # redraws done so far = the `draw` word of the next stream
        self._stale = None              # (host training array, n_bc, draw) whose host copy lags behind the device rows

    # ---- deepxde.callbacks.Callback protocol -------------------------------------------------------------------
    def set_model(self, model):
        self.model = model

    def init(self): pass
    def on_epoch_begin(self): pass
    def on_batch_begin(self): pass
    def on_batch_end(self): pass
    def on_predict_begin(self): pass
    def on_predict_end(self): pass

    def on_train_begin(self):
        core = getattr(self.model, "fused_core", None)
        if core is None:
            raise capi.FusedOpError("DevicePointResampler needs a model patched by enable_fused()")
        data = self.model.data
        if getattr(data, "exclusions", None) is not None:
            raise capi.FusedOpError("device resampler: PDE(exclusions=...) filters the drawn points on the host; not supported")
        if int(getattr(data, "num_domain", 0)) <= 0:
            raise capi.FusedOpError("device resampler: the problem has no domain points (num_domain == 0)")
        if self.region is None:
            self.region = geometry_region(data.geom)
        d_points = core.op.spec.embed.shape[0] if core.op.spec.embed is not None else core.op.desc.d
        if self.region.d != d_points:
            raise capi.FusedOpError(f"device resampler: region has {self.region.d} coordinates, the problem's points {d_points}")
        self.num_bcs_initial = list(data.num_bcs)

    def on_epoch_end(self):
        self.epochs_since_last_resample += 1
        if self.epochs_since_last_resample < self.period:
            return
        self.epochs_since_last_resample = 0
        self.resample()

    def on_train_end(self):
        self.sync_host()

    # ---- the redraw ------------------------------------------------------------------------------------------------
    def resample(self):
        """Redraw the ``num_domain`` interior rows of the staged training set in place (one launch + the aux columns)."""
        data, core = self.model.data, self.model.fused_core
        n_bc = int(np.sum(data.num_bcs))
        core.redraw_domain_rows(data.train_x, n_bc, int(data.num_domain), self.region, self.seed, self.draws, on_evict=self.sync_host)
        self._stale = (data.train_x, n_bc, self.draws)
        self.draws += 1

    def sync_host(self):
        """Bring the host arrays (``data.train_x`` = ``train_state.X_train``, ``data.train_x_all``) up to date with the
        device rows, in place -- the arrays keep their identity, so nothing is re-uploaded afterwards."""
        if self._stale is None:
            return
        data = self.model.data
        host, n_bc, draw = self._stale
        n_dom = int(data.num_domain)
        dev = self.model.fused_core.op.device
        rows = sample_region(self.region, n_dom, self.seed, draw, 0, device=dev).cpu().numpy()
        host[host.shape[0] - n_dom:, :self.region.d] = rows
        all_rows = getattr(data, "train_x_all", None)
        if all_rows is not None and all_rows is not host:
            all_rows[all_rows.shape[0] - n_dom:, :self.region.d] = rows
        self._stale = None
