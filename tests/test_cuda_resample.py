"""Device point sampler on the GPU: ``pinn_sample_points`` through the C ABI is bit-identical to
oracle/sample_ref.py on the reference's own geometries (tests/golden/regions.npz), at 16 Mi rows through size-independent
properties, and the resampling callback keeps a patched model's device rows, aux columns and host arrays consistent."""
import json
import os

import numpy as np
import pytest
import torch

from test_cuda_solver import DirichletBC, _StubData, _StubModel, _net

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _regions():
    z = np.load(os.path.join(ROOT, "tests", "golden", "regions.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


@pytest.mark.parametrize("name", ["Poisson2D_Classic", "Heat2D_ComplexGeometry", "NS2D_Classic", "NS2D_BackStep", "Poisson3D_ComplexGeometry",
                                  "PoissonBoltzmann2D", "Burgers1D", "GrayScottEquation", "HeatND", "PoissonND"])
def test_kernel_is_bit_identical_to_the_oracle_on_reference_geometries(name):
    from oracle.sample_ref import in_holes, outside_ball, sample_points
    from pinnacle_b200.resample import Region, sample_region
    _, meta = _regions()
    m = meta[name]
    region = Region(m["lo"], m["hi"], [tuple(h) for h in m["holes"]], None if m["keep"] is None else (m["keep"][0], m["keep"][1]))
    for n, key, draw, first in ((10007, 5, 0, 0), (777, (1 << 63) + 12345, (1 << 24) - 1, (1 << 32) - 300), (1, 0, 3, 1 << 40), (256, 1, 1, 0)):
        got = sample_region(region, n, key, draw, first, device="cuda:0").cpu().numpy()
        want = sample_points(n, region.lo, region.hi, region.holes, key=key, draw=draw, first_row=first, keep=region.keep)
        assert np.array_equal(got, want), (name, n, key, draw, first, np.flatnonzero((got != want).any(1))[:5])
        assert not in_holes(got, region.holes).any() and not outside_ball(got, region.keep).any()
    assert sample_region(region, 0, device="cuda:0").shape == (0, region.d)


def test_strided_rows_six_coordinates_and_refusals():
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import capi
    from pinnacle_b200.kinds import CRegion
    from pinnacle_b200.resample import Region
    dev = torch.device("cuda:0")
    # rows of 4 floats of which the first 2 are coordinates (the embedded gPINN layout): columns 2, 3 are left alone
    reg = Region((-1, 0), (1, 2), [(0, (0.0, 1.0), (0.5,))])
    x = torch.full((1000, 4), 7.0, device=dev)
    capi.sample_points(x, reg.to_c(), 3, 1, 100)
    assert np.array_equal(x[:, :2].cpu().numpy(), sample_points(1000, reg.lo, reg.hi, reg.holes, key=3, draw=1, first_row=100))
    assert bool((x[:, 2:] == 7.0).all())
    # a row slice of a larger array: the rows outside stay
    y = torch.zeros((600, 3), device=dev)
    reg3 = Region((0, 0, 0), (1, 1, 5))
    capi.sample_points(y[100:357], reg3.to_c(), 9, 0, 0)
    assert bool((y[:100] == 0).all()) and bool((y[357:] == 0).all())
    assert np.array_equal(y[100:357].cpu().numpy(), sample_points(257, reg3.lo, reg3.hi, key=9))
    # six coordinates = two Philox blocks per attempt, hole over the first three
    reg6 = Region((0,) * 6, (1, 2, 3, 4, 5, 6), [(1, (0.2, 0.2, 0.2), (0.8, 1.8, 2.8))])
    z = torch.empty((5000, 6), device=dev)
    capi.sample_points(z, reg6.to_c(), 1, 2, 3)
    assert np.array_equal(z.cpu().numpy(), sample_points(5000, reg6.lo, reg6.hi, reg6.holes, key=1, draw=2, first_row=3))
    # refusals come back through pinn_last_error
    bad = reg.to_c()
    bad.d = 0
    with pytest.raises(capi.FusedOpError, match="PINN_MAX_DIM"):
        capi.sample_points(x, bad, 0, 0)
    bad = reg.to_c()
    bad.holes[0].n_dims = 3
    with pytest.raises(capi.FusedOpError, match="hole"):
        capi.sample_points(x, bad, 0, 0)
    with pytest.raises(capi.FusedOpError, match="draw"):
        capi.sample_points(x, reg.to_c(), 0, 1 << 24)
    bad = CRegion()
    bad.d, bad.lo[0], bad.hi[0] = 1, 1.0, 0.0
    with pytest.raises(capi.FusedOpError, match="empty"):
        capi.sample_points(x, bad, 0, 0)
    with pytest.raises(capi.FusedOpError, match="CUDA tensor"):
        capi.sample_points(torch.empty(4, 2), reg.to_c(), 0, 0)


def test_sixteen_million_rows_shard_consistency_moments_and_rate():
    """BASELINE.json's largest point set (16 Mi rows x 3): two half-range launches equal the whole (a rank draws the rows one
    rank would), every row is outside the holes, the box-uniform moments hold, and the launch runs at HBM-write speed."""
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import capi
    from pinnacle_b200.resample import Region
    dev = torch.device("cuda:0")
    n = 1 << 24
    _, meta = _regions()
    m = meta["Heat2D_ComplexGeometry"]                             # 17 disks
    reg = Region(m["lo"], m["hi"], [tuple(h) for h in m["holes"]])
    c = reg.to_c()
    whole = torch.empty((n, 3), device=dev)
    capi.sample_points(whole, c, 77, 5, 0)
    halves = torch.empty((n, 3), device=dev)
    cut = 7_000_001
    capi.sample_points(halves[:cut], c, 77, 5, 0)
    capi.sample_points(halves[cut:], c, 77, 5, cut)
    assert torch.equal(whole, halves)
    # spot rows against the oracle
    for first in (0, cut - 3, n - 1000):
        assert np.array_equal(whole[first:first + 1000].cpu().numpy(), sample_points(min(1000, n - first), reg.lo, reg.hi, reg.holes, key=77, draw=5, first_row=first))
    # no row in a hole (the rule evaluated with torch on the device, same float32 operations)
    for s, a, b in reg.holes:
        d2 = torch.zeros(n, device=dev)
        for j, aj in enumerate(a):
            dj = whole[:, j] - np.float32(aj)
            d2 = d2 + dj * dj
        assert not bool((d2 <= np.float32(b[0]) * np.float32(b[0])).any())
    lo, hi = torch.tensor(reg.lo, device=dev), torch.tensor(reg.hi, device=dev)
    assert bool((whole >= lo).all()) and bool((whole <= hi).all())
    t = whole[:, 2].double()
    assert abs(float(t.mean()) - 0.5 * (reg.lo[2] + reg.hi[2])) < 5 * (reg.hi[2] - reg.lo[2]) / np.sqrt(12 * n)
    # timing: CUDA events on the launching stream, L2 flushed by the 192 MB the launch itself writes
    plain = Region(reg.lo, reg.hi).to_c()
    times = {}
    for tag, cc in (("17 holes", c), ("no holes", plain)):
        for _ in range(3):
            capi.sample_points(whole, cc, 1, 0, 0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(10):
            capi.sample_points(whole, cc, 1, i, 0)
        ev[1].record()
        torch.cuda.synchronize()
        times[tag] = ev[0].elapsed_time(ev[1]) / 10
    gbs = {k: n * 12 / (v * 1e-3) / 1e9 for k, v in times.items()}
    print(f"\npinn_sample_points, 16 Mi x 3 rows: {times} ms per launch = {gbs} GB/s written")
    assert times["no holes"] < 1.0 and times["17 holes"] < 2.0, times       # 201 MB; the host path is ~0.3 s of numpy + a 201 MB upload


def _patched(spec_name, n_rows, n_domain, nb=256):
    from pinnacle_b200 import problems
    from pinnacle_b200.solver import enable_fused
    dev = torch.device("cuda:0")
    spec = problems.make(spec_name)
    x = spec.sample(n_rows, seed=3)
    xb = spec.sample(nb, seed=4)
    vals = np.random.RandomState(1).standard_normal((nb, 1)).astype(np.float32)
    train_x = np.vstack([xb, x]).astype(np.float32)
    net = _net(spec.desc, dev, 23)
    data = _StubData(lambda x_, u_: None, [DirichletBC(0, vals)], [nb], train_x)
    data.num_domain, data.exclusions, data.train_x_all = n_domain, None, train_x[nb:].copy()
    model = _StubModel(data, net, torch.optim.Adam(net.parameters(), 1e-3), np.ones(2))
    enable_fused(model, spec=spec)
    return model, spec, net, train_x, nb


@pytest.mark.parametrize("spec_name", ["Poisson2D_Classic", "PoissonBoltzmann2D", "Heat2D_LongTime"])
def test_callback_redraw_equals_an_upload_of_the_same_rows(spec_name):
    """After a device redraw the patched model's losses and gradients equal, bit for bit, those of a fresh model that was
    handed the oracle's rows through the ordinary upload path (same kernels, same rows, same aux columns)."""
    from oracle.sample_ref import sample_points
    from pinnacle_b200.resample import DevicePointResampler, Region
    n_rows, n_dom = 40_000, 30_000
    model, spec, net, train_x, nb = _patched(spec_name, n_rows, n_dom)
    region = Region(spec.bbox[0::2], spec.bbox[1::2], [(0, tuple(0.5 * (l + h) for l, h in zip(spec.bbox[0:4:2], spec.bbox[1:4:2])), (0.2,))])
    cb = DevicePointResampler(period=2, seed=99, region=region)
    cb.set_model(model)
    cb.on_train_begin()
    host_before = train_x.copy()
    _, l0 = model.outputs_losses_train(train_x, None)
    uploads = model.fused_core.op.stats["uploads"]
    for _ in range(4):
        cb.on_epoch_end()
    assert cb.draws == 2
    _, l1 = model.outputs_losses_train(train_x, None)
    net.zero_grad()
    l1.sum().backward()
    g1 = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    assert model.fused_core.op.stats["uploads"] == uploads               # nothing was uploaded
    assert np.array_equal(train_x, host_before)                           # the host copy lags ...
    assert float((l1[0] - l0[0]).abs()) > 0
    cb.on_train_end()                                                     # ... until the end of train()
    want = sample_points(n_dom, region.lo, region.hi, region.holes, key=99, draw=1)
    assert np.array_equal(train_x[-n_dom:], want) and np.array_equal(train_x[:-n_dom], host_before[:-n_dom])
    assert np.array_equal(model.data.train_x_all[-n_dom:], want)
    _, l1b = model.outputs_losses_train(train_x, None)                    # same array identity: still no upload, same rows
    assert model.fused_core.op.stats["uploads"] == uploads and torch.allclose(l1b, l1, rtol=1e-6, atol=0)
    # a fresh model fed the same rows from the host
    model2, _, net2, _, _ = _patched(spec_name, n_rows, n_dom)
    fresh = train_x.copy()
    _, l2 = model2.outputs_losses_train(fresh, None)
    net2.zero_grad()
    l2.sum().backward()
    g2 = torch.cat([p.grad.reshape(-1) for p in net2.parameters()])
    assert torch.allclose(l2, l1, rtol=1e-6, atol=0), (l2, l1)                # loss partials are summed with atomics too
    assert float((g2 - g1).norm()) <= 1e-6 * float(g1.norm())             # gradient partials are summed with atomics: order varies


def test_redraw_under_a_captured_step_graph():
    """The redraw writes in place, so a CUDA graph captured over the training step keeps reading the live rows: after a
    redraw the replayed step's losses are those of an eager solver that was handed the oracle's rows from the host."""
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import problems
    from pinnacle_b200.optim import FlatAdam
    from pinnacle_b200.resample import Region
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec = problems.make("Heat2D_Multiscale")
    n = 50_000
    x = spec.sample(n, seed=1)
    region = Region(spec.bbox[0::2], spec.bbox[1::2])
    net = _net(spec.desc, dev, 5)
    solver = FusedPDESolver(spec, net, x).compile(FlatAdam(net.parameters(), lr=0.0))      # lr 0: the weights stay, only the rows change
    solver.capture_step()
    solver.train_step()
    before = solver.opt.losses.detach().clone()
    solver.redraw_points(region, n_domain=n - 1000, key=4, draw=0)
    solver.train_step()
    after = solver.opt.losses.detach().clone()
    assert solver._graph is not None and not torch.equal(before, after)
    rows = np.vstack([x[:1000], sample_points(n - 1000, region.lo, region.hi, key=4, draw=0)]).astype(np.float32)
    net2 = _net(spec.desc, dev, 5)
    eager = FusedPDESolver(spec, net2, rows).compile(FlatAdam(net2.parameters(), lr=0.0))
    eager.train_step()
    assert torch.allclose(eager.opt.losses.detach(), after, rtol=1e-6, atol=0), (eager.opt.losses, after)
