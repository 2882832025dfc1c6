"""Device point sampler, CPU side: the oracle (oracle/sample_ref.py) against the Philox known-answer vectors and the
reference-generated region fixture (tests/golden/regions.npz), the geometry -> region mapping, the struct layout, and the
resampling callback's host logic with the oracle standing in for the kernel."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import refshim
from util import oracle_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


def regions():
    z = np.load(os.path.join(ROOT, "tests", "golden", "regions.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    from oracle.sample_ref import philox4x32_10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_rows_are_a_function_of_their_index_only():
    from oracle.sample_ref import HOLE_BALL, sample_points
    holes = [(HOLE_BALL, (0.3, 0.3), (0.25,))]
    a = sample_points(1000, (0, 0, 0), (1, 1, 5), holes, key=7, draw=3)
    b = sample_points(300, (0, 0, 0), (1, 1, 5), holes, key=7, draw=3, first_row=450)
    assert a.dtype == np.float32 and np.array_equal(a[450:750], b)           # a shard draws the rows one rank would
    assert not np.array_equal(a, sample_points(1000, (0, 0, 0), (1, 1, 5), holes, key=7, draw=4))
    assert not np.array_equal(a, sample_points(1000, (0, 0, 0), (1, 1, 5), holes, key=8, draw=3))
    big = sample_points(4, (0, 0), (1, 1), key=1, first_row=(1 << 32) + 5)   # 64-bit row counter
    assert not np.array_equal(big, sample_points(4, (0, 0), (1, 1), key=1, first_row=5))


def test_distribution_is_uniform_on_box_minus_holes():
    from oracle.sample_ref import HOLE_BALL, HOLE_BOX, in_holes, sample_points
    holes = [(HOLE_BALL, (0.5, 0.5), (0.3,)), (HOLE_BOX, (-1.0, -1.0), (-0.5, 0.0))]
    n = 200_000
    p = sample_points(n, (-1, -1, 0), (1, 1, 2), holes, key=11)
    assert (p >= np.array([-1, -1, 0], np.float32)).all() and (p < np.array([1, 1, 2], np.float32)).all()
    assert not in_holes(p, holes).any()
    free = 4.0 - np.pi * 0.09 - 0.5
    # the time coordinate is untouched by the rejection; a sub-box away from the holes carries its share of the free area
    assert abs(p[:, 2].mean() - 1.0) < 4 * (2 / np.sqrt(12)) / np.sqrt(n)
    frac = ((p[:, 0] > -0.4) & (p[:, 0] < 0.1) & (p[:, 1] < -0.2)).mean()
    want = 0.5 * 0.8 / free
    assert abs(frac - want) < 4 * np.sqrt(want * (1 - want) / n)
    # 6 coordinates take two Philox blocks; the second block is independent of the first
    q = sample_points(50_000, (0,) * 6, (1,) * 6, key=3)
    c = np.corrcoef(q.T)
    assert np.abs(c - np.eye(6)).max() < 0.03 and np.abs(q.mean(0) - 0.5).max() < 0.01


def test_ball_domain_is_uniform():
    """HeatND's domain: the unit 5-ball times [0, 1], drawn by rejection from the bounding cube (16 % accepted per attempt,
    so rows need up to dozens of attempts).  Uniform on the ball: E|x|^2 = d/(d+2), P(|x| < r) = r^d; t untouched."""
    from oracle.sample_ref import outside_ball, sample_points
    n = 100_000
    keep = ((0.0,) * 5, 1.0)
    p = sample_points(n, (-1,) * 5 + (0,), (1,) * 5 + (1,), key=2, keep=keep)
    assert not outside_ball(p, keep).any()
    r2 = (p[:, :5].astype(np.float64) ** 2).sum(1)
    assert abs(r2.mean() - 5 / 7) < 5 * np.sqrt(np.var(r2) / n)
    assert abs((r2 < 0.8 ** 2).mean() - 0.8 ** 5) < 5 * np.sqrt(0.8 ** 5 * (1 - 0.8 ** 5) / n)
    assert np.abs(p[:, :5].mean(0)).max() < 5 * np.sqrt(1 / 7 / n) and abs(p[:, 5].mean() - 0.5) < 5 / np.sqrt(12 * n)


@pytest.mark.parametrize("name", ["Poisson2D_Classic", "Heat2D_ComplexGeometry", "NS2D_Classic", "NS2D_BackStep", "Poisson3D_ComplexGeometry",
                                  "PoissonBoltzmann2D", "Burgers1D", "HeatND", "PoissonND"])
def test_hole_rule_equals_the_reference_geometry_inside(name):
    """4096 probe points per class (some pulled onto the hole boundaries) with the verdicts of the reference's own
    ``geom.inside`` (tests/golden/make_region_golden.py): the region's point-in-hole rule must agree, the reference's own draw
    must be accepted, and our draw must be accepted by the stored verdict rule."""
    from oracle.sample_ref import in_holes, outside_ball, sample_points
    z, meta = regions()
    m = meta[name]
    holes = [(s, a, b) for s, a, b in m["holes"]]
    keep = None if m["keep"] is None else (m["keep"][0], m["keep"][1])
    rejected = lambda p: in_holes(p, holes) | outside_ball(p, keep)     # noqa: E731
    probe, inside = z[name + "_probe"], z[name + "_inside"]
    ours = ~rejected(probe)
    # float32 (ours) against float64 norms (numpy's) can differ on points within rounding of a ball's boundary
    balls = [(a, b[0]) for s, a, b in holes if s == 0] + ([keep] if keep else [])
    disagree = np.flatnonzero(ours != inside)
    for i in disagree:
        dist = min(abs(np.linalg.norm(probe[i, :len(a)].astype(np.float64) - np.asarray(a)) - r) for a, r in balls)
        assert dist < 1e-6, (name, i, probe[i], dist)
    assert len(disagree) <= 2
    refdraw = z[name + "_refdraw"]
    bad = np.flatnonzero(rejected(refdraw))                               # the reference's own draw; radius**(1/d) can land ON the sphere
    for i in bad:
        assert keep and abs(np.linalg.norm(refdraw[i, :len(keep[0])].astype(np.float64) - np.asarray(keep[0])) - keep[1]) < 1e-6
    own = sample_points(5000, m["lo"], m["hi"], holes, key=5, keep=keep)
    assert not rejected(own).any()
    assert (own >= np.asarray(m["lo"], np.float32)).all() and (own <= np.asarray(m["hi"], np.float32)).all()
    # acceptance: the share of box-uniform probes the reference keeps = the share of first attempts we keep
    first = sample_points(4096, m["lo"], m["hi"], [], key=5)
    n_pulled = 24 * len(holes)
    body = slice(n_pulled, 4096 - (64 if keep else 0))
    assert abs((~rejected(first)).mean() - inside[body].mean()) < 0.04


class Interval:
    def __init__(self, l, r): self.l, self.r = l, r


class TimeDomain(Interval):
    def __init__(self, t0, t1):
        super().__init__(t0, t1)
        self.t0, self.t1 = t0, t1


class Rectangle:
    def __init__(self, xmin, xmax): self.xmin, self.xmax = np.array(xmin), np.array(xmax)


class Cuboid(Rectangle): pass


class Disk:
    def __init__(self, c, r): self.center, self.radius = np.array(c), r


class Sphere(Disk): pass


class Hypersphere(Disk): pass


class CSGDifference:
    def __init__(self, a, b): self.geom1, self.geom2 = a, b


class CSGUnion(CSGDifference): pass


class CSGMultiDifference:
    def __init__(self, a, bs): self.geom1, self.geom2_list = a, bs


class GeometryXTime:
    def __init__(self, g, t): self.geometry, self.timedomain = g, t


class Polygon: pass


def test_geometry_to_region_mapping():
    from pinnacle_b200 import capi
    from pinnacle_b200.kinds import HOLE_BALL, HOLE_BOX
    from pinnacle_b200.resample import Region, geometry_region
    r = geometry_region(GeometryXTime(Interval(-1, 1), TimeDomain(0, 2)))
    assert (r.lo, r.hi, r.holes) == ((-1.0, 0.0), (1.0, 2.0), [])
    g = CSGDifference(CSGDifference(Rectangle([0, 0], [4, 2]), Disk([1, 1], 0.5)), CSGUnion(Rectangle([-1, 1], [2, 3]), Disk([3, 1], 0.25)))
    r = geometry_region(GeometryXTime(g, TimeDomain(0, 5)))
    assert r.d == 3 and r.lo == (0.0, 0.0, 0.0) and r.hi == (4.0, 2.0, 5.0)
    assert r.holes == [(HOLE_BALL, (1.0, 1.0), (0.5,)), (HOLE_BOX, (-1.0, 1.0), (2.0, 3.0)), (HOLE_BALL, (3.0, 1.0), (0.25,))]
    r = geometry_region(CSGMultiDifference(Cuboid([0, 0, 0], [1, 1, 1]), [Sphere([0.5, 0.5, 0.5], 0.2), Sphere([0.1, 0.1, 0.1], 0.05)]))
    assert r.d == 3 and len(r.holes) == 2 and r.holes[1] == (HOLE_BALL, (0.1, 0.1, 0.1), (0.05,))
    c = r.to_c()
    assert (c.d, c.n_holes, c.holes[0].n_dims) == (3, 2, 3) and abs(c.holes[0].b[0] - 0.2) < 1e-7 and c.hi[2] == 1.0
    assert 0.03 < r.volume_fraction_bound() < 0.04
    r = geometry_region(GeometryXTime(Hypersphere([0] * 5, 1), TimeDomain(0, 2)))          # HeatND: the ball inside its bounding cube
    assert r.d == 6 and r.lo == (-1.0,) * 5 + (0.0,) and r.hi == (1.0,) * 5 + (2.0,) and r.keep == ((0.0,) * 5, 1.0) and r.holes == []
    c = r.to_c()
    assert (c.keep_dims, c.keep_radius, c.n_holes) == (5, 1.0, 0) and 0.83 < r.volume_fraction_bound() < 0.84
    r = geometry_region(CSGDifference(Disk([1, 2], 3), Rectangle([0, 0], [1, 1])))
    assert r.lo == (-2.0, -1.0) and r.keep == ((1.0, 2.0), 3.0) and len(r.holes) == 1
    for bad in (Polygon(), CSGDifference(Rectangle([0, 0], [1, 1]), Polygon()), CSGUnion(Rectangle([0, 0], [1, 1]), Disk([0, 0], 1)),
                CSGDifference(Cuboid([0] * 4, [1] * 4), Hypersphere([0] * 4, 0.1))):
        with pytest.raises(capi.FusedOpError):
            geometry_region(bad)
    with pytest.raises(ValueError):
        Region((0, 0), (1, -1))
    with pytest.raises(ValueError):
        Region((0,), (1,), [(HOLE_BALL, (0.5, 0.5), (0.1,))])       # a hole over more coordinates than the region has


def test_region_struct_layout_matches_header():
    from pinnacle_b200 import kinds
    txt = open(os.path.join(ROOT, "include", "pinn_b200.h")).read()
    assert int(re.search(r"#define PINN_MAX_HOLES (\d+)", txt).group(1)) == kinds.PINN_MAX_HOLES
    assert ctypes.sizeof(kinds.CHole) == 8 + 4 * 3 + 4 * 3
    assert ctypes.sizeof(kinds.CRegion) == 8 + 2 * 4 * kinds.PINN_MAX_DIM + 8 + 4 * kinds.PINN_MAX_DIM + kinds.PINN_MAX_HOLES * ctypes.sizeof(kinds.CHole)
    assert kinds.CRegion.holes.offset == 8 + 3 * 4 * kinds.PINN_MAX_DIM + 8 and kinds.CRegion.keep_center.offset == 8 + 2 * 4 * kinds.PINN_MAX_DIM + 8
    assert re.search(r"PINN_HOLE_BALL = (\d+), PINN_HOLE_BOX = (\d+)", txt).groups() == (str(kinds.HOLE_BALL), str(kinds.HOLE_BOX))


def test_sampler_refuses_cpu_tensors():
    """No CPU implementation behind the product call: a host tensor is an error, not a fallback."""
    from pinnacle_b200 import capi
    from pinnacle_b200.resample import Region, sample_region
    with pytest.raises(capi.FusedOpError, match="CUDA tensor"):
        sample_region(Region((0, 0), (1, 1)), 8, out=torch.empty(8, 2))


def _core(spec, cache_points=True):
    from pinnacle_b200 import fused
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.solver import _StepCore
    torch.manual_seed(0)
    net = FNN([spec.net_d] + [100] * 5 + [spec.desc.m])
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        return _StepCore(spec, net, "auto", "cpu", None, cache_points), net
    finally:
        fused.FusedPDEResidual.__init__ = orig


def test_redraw_replaces_the_last_rows_in_place_and_refreshes_aux():
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import problems
    from pinnacle_b200.resample import Region
    spec = problems.make("PoissonBoltzmann2D")                    # has a per-point source column
    region = Region(spec.bbox[0::2], spec.bbox[1::2], [(0, (0.5, 0.5), (0.3,))])
    with oracle_backend():
        core, net = _core(spec)
        X = np.concatenate([spec.sample(10, seed=1), spec.sample(90, seed=2)]).astype(np.float32)     # 10 BC rows + 90 residual rows
        core.stage(X, 10)
        x_dev, aux_dev = core.op.x, core.op.aux
        before, aux_before = x_dev.clone(), aux_dev.clone()
        gen = core.generation
        core.redraw_domain_rows(X, 10, 64, region, key=9, draw=2)
        assert core.op.x is x_dev and core.op.aux is aux_dev and core.generation == gen           # in place: nothing restaged
        assert torch.equal(x_dev[:26], before[:26])                                               # boundary / initial rows of the residual set stay
        want = sample_points(64, region.lo, region.hi, region.holes, key=9, draw=2)
        assert np.array_equal(x_dev[26:].numpy(), want)
        assert torch.equal(aux_dev[:26], aux_before[:26]) and not torch.equal(aux_dev[26:], aux_before[26:])
        assert torch.allclose(aux_dev, spec.aux(x_dev))
        assert core.op.stats["device_redraws"] == 1
        with pytest.raises(Exception):
            core.redraw_domain_rows(X, 10, 91, region, key=9, draw=2)


class _FakeData:
    def __init__(self, X, n_bc, n_domain, geom):
        self.train_x, self.num_bcs, self.num_domain, self.geom = X, [n_bc], n_domain, geom
        self.train_x_all = X[n_bc:].copy()
        self.exclusions = None


class _FakeModel:
    pass


def test_callback_schedule_and_host_sync():
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import capi, problems
    from pinnacle_b200.resample import DevicePointResampler
    spec = problems.make("Heat2D_Multiscale")
    geom = GeometryXTime(Rectangle([0, 0], [1, 1]), TimeDomain(0, 5))
    with oracle_backend():
        core, net = _core(spec)
        X = spec.sample(120, seed=3).astype(np.float32)
        X0 = X.copy()
        model = _FakeModel()
        model.data, model.fused_core = _FakeData(X, 20, 80, geom), core
        core.stage(X, 20)
        cb = DevicePointResampler(period=3, seed=42)
        cb.set_model(model)
        cb.on_train_begin()
        assert cb.region.d == 3 and cb.region.hi == (1.0, 1.0, 5.0)
        for _ in range(2):
            cb.on_epoch_end()
        assert cb.draws == 0 and torch.equal(core.op.x, torch.from_numpy(X0[20:]))
        cb.on_epoch_end()
        assert cb.draws == 1
        for _ in range(3):
            cb.on_epoch_end()
        assert cb.draws == 2
        want = sample_points(80, (0, 0, 0), (1, 1, 5), key=42, draw=1)
        assert np.array_equal(core.op.x[20:].numpy(), want) and np.array_equal(core.op.x[:20].numpy(), X0[20:40])
        # (on a GPU the host copy lags until sync_host(): tests/test_cuda_resample.py; here the CPU "device" rows alias X)
        cb.on_train_end()
        assert np.array_equal(X[40:], want) and np.array_equal(X[:40], X0[:40])
        assert np.array_equal(model.data.train_x_all[20:], want)
        key_before, gen = core._key, core.generation
        core.stage(X, 20)                                             # same array identity: no re-upload after the sync
        assert core._key == key_before and core.generation == gen and len(core._staged) == 1
        # refusals
        model.data.exclusions = np.zeros((1, 3))
        with pytest.raises(capi.FusedOpError, match="exclusions"):
            cb.on_train_begin()
        model.data.exclusions = None
        model.data.geom = GeometryXTime(Polygon(), TimeDomain(0, 5))
        cb2 = DevicePointResampler(period=3)
        cb2.set_model(model)
        with pytest.raises(capi.FusedOpError, match="Polygon"):
            cb2.on_train_begin()
        model.data.geom = Rectangle([0, 0], [1, 1])
        with pytest.raises(capi.FusedOpError, match="coordinates"):
            cb2.on_train_begin()
        del model.fused_core
        with pytest.raises(capi.FusedOpError, match="enable_fused"):
            cb2.on_train_begin()


@needs_ref
@pytest.mark.parametrize("name", ["Poisson2D_Classic", "Burgers1D", "Heat2D_ComplexGeometry", "HeatND"])
def test_resampler_on_the_reference_model(name):
    """The real dde.Model, patched, trains with the device resampler as a callback: every redraw lands inside the
    reference geometry, the number of rows and the boundary rows never change, host arrays are current after train()."""
    dde, _ = refshim.load()
    from pinnacle_b200 import fused
    from pinnacle_b200.resample import DevicePointResampler
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    dde.config.set_random_seed(0)
    pde = refshim.construct(name)
    pde.training_points(domain=96, boundary=400, **({"initial": 32} if hasattr(pde, "num_initial_points") else {}))
    pde.num_test_points = 64
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
    model = pde.create_model(net)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            X_before = model.data.train_x.copy()
            cb = DevicePointResampler(period=2, seed=1)
            model.train(iterations=5, display_every=5, callbacks=[cb])
            assert cb.draws == 2
            X = model.data.train_x
            n_dom = model.data.num_domain
            assert X.shape == X_before.shape and np.array_equal(X[:-n_dom], X_before[:-n_dom])
            assert not np.array_equal(X[-n_dom:], X_before[-n_dom:])
            geom = model.data.geom
            spatial = geom.geometry if type(geom).__name__ == "GeometryXTime" else geom
            pts = X[-n_dom:]
            assert spatial.inside(pts[:, :spatial.dim]).all()
            assert np.array_equal(model.data.train_x_all[-n_dom:], pts)
            assert np.isfinite(model.train_state.loss_train).all()
    finally:
        fused.FusedPDEResidual.__init__ = orig


@needs_ref
def test_resampler_and_rar_hand_over_through_the_host_arrays():
    """RAR (evaluate.fused_rar_wrapper, the reference's schedule) between training intervals reads ``train_state.X_train`` on the
    host and prepends anchors with ``data.add_anchors``; the device resampler must have brought that array up to date at the
    end of the interval (on_train_end), and keeps redrawing the -- still last -- interior rows of the extended set afterwards."""
    dde, _ = refshim.load()
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import fused
    from pinnacle_b200.evaluate import fused_rar_wrapper
    from pinnacle_b200.resample import DevicePointResampler
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    dde.config.set_random_seed(0)
    pde = refshim.construct("Burgers1D")
    pde.training_points(domain=64, boundary=32, initial=32)
    pde.num_test_points = 64
    net = dde.nn.FNN([2] + [100] * 5 + [1], "tanh", "Glorot normal").float()
    model = pde.create_model(net)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            n_all = model.data.train_x_all.shape[0]
            cb = DevicePointResampler(period=2, seed=5)
            train = fused_rar_wrapper(pde, model, {"interval": 4, "count": 3})
            train(iterations=12, display_every=4, callbacks=[cb])
            # 3 intervals of 4 steps -> 2 redraws each; 2 RAR rounds of 3 anchors in between
            assert cb.draws == 6
            data = model.data
            assert data.train_x_all.shape[0] == n_all + 6 and data.anchors.shape[0] == 6
            reg = cb.region
            want = sample_points(64, reg.lo, reg.hi, reg.holes, key=5, draw=5)
            assert np.array_equal(data.train_x_all[-64:], want) and np.array_equal(data.train_x[-64:], want)
            # the anchors of the second RAR round were picked among rows that include the 4th redraw's interior points
            prev = sample_points(64, reg.lo, reg.hi, reg.holes, key=5, draw=3)
            pool = np.vstack([data.train_x_all[6:-64], prev])
            for a in data.anchors[:3]:
                assert (np.abs(pool - a).max(axis=1) == 0).any() or (np.abs(data.anchors[3:] - a).max(axis=1) == 0).any()
            assert np.isfinite(model.train_state.loss_train).all()
    finally:
        fused.FusedPDEResidual.__init__ = orig


def test_reference_free_solver_redraws_its_points():
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import fused, problems
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.resample import Region
    from pinnacle_b200.solver import FusedPDESolver
    spec = problems.make("Heat2D_Multiscale")
    region = Region(spec.bbox[0::2], spec.bbox[1::2])
    torch.manual_seed(0)
    net = FNN([3] + [100] * 5 + [1])
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            x = spec.sample(60, seed=1).astype(np.float32).copy()
            solver = FusedPDESolver(spec, net, x, device="cpu").compile("adam")
            l0 = solver.outputs_losses_train()[1].detach().clone()
            solver.redraw_points(region, key=3, draw=1)
            assert np.array_equal(solver.core.op.x.numpy(), sample_points(60, region.lo, region.hi, key=3, draw=1))
            l1 = solver.outputs_losses_train()[1].detach()
            assert not torch.equal(l0, l1)
            solver.redraw_points(region, n_domain=20, key=3, draw=2)
            assert np.array_equal(solver.core.op.x[40:].numpy(), sample_points(20, region.lo, region.hi, key=3, draw=2))
            assert np.array_equal(solver.core.op.x[:40].numpy(), sample_points(60, region.lo, region.hi, key=3, draw=1)[:40])
            solver.train_step()
    finally:
        fused.FusedPDEResidual.__init__ = orig


def test_host_copy_is_refreshed_before_a_redrawn_set_leaves_the_device():
    """The resident point sets are an LRU of three.  A training set whose interior rows were redrawn on the device must not be
    dropped while its host array still holds the old rows -- a later re-upload would silently bring them back."""
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import problems
    from pinnacle_b200.resample import DevicePointResampler
    spec = problems.make("Heat2D_Multiscale")
    geom = GeometryXTime(Rectangle([0, 0], [1, 1]), TimeDomain(0, 5))
    with oracle_backend():
        core, net = _core(spec)
        X = spec.sample(100, seed=3).astype(np.float32)
        X0 = X.copy()
        model = _FakeModel()
        model.data, model.fused_core = _FakeData(X, 10, 70, geom), core
        # on the CPU "device" the staged rows alias the host array; stage a copy so that the host array really lags
        core.stage(X, 10)
        ent = core._staged[core._key]
        ent["x"] = ent["x"].clone()
        core.op.x = ent["x"]
        cb = DevicePointResampler(period=1, seed=8)
        cb.set_model(model)
        cb.on_train_begin()
        cb.on_epoch_end()
        assert np.array_equal(X, X0) and cb._stale is not None
        for s in range(3):                                          # three other point sets push the training set out
            core.stage(spec.sample(40, seed=20 + s).astype(np.float32), 5)
        assert len(core._staged) == core.MAX_STAGED and cb._stale is None
        want = sample_points(70, (0, 0, 0), (1, 1, 5), key=8, draw=0)
        assert np.array_equal(X[30:], want) and np.array_equal(X[:30], X0[:30])
        core.stage(X, 10)                                           # re-upload: the redrawn rows come back, not the old ones
        assert np.array_equal(core.op.x[20:].numpy(), want)
