"""N>1 host path on CPU: world_size-2 gloo.  The sharding + all-reduce plumbing of the fused op is exercised
with the oracle standing in for the C ABI (explicitly injected by the test; the product never does this)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, policy, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pinnacle_b200 import problems
        from pinnacle_b200.fused import FusedPDEResidual, shard_bounds
        from pinnacle_b200.nn import FNN
        from util import oracle_backend
        torch.manual_seed(0)
        torch.set_num_threads(2)
        spec = problems.make(name)
        desc = spec.desc
        net = FNN([desc.d] + [100] * 5 + [desc.m])
        with torch.no_grad():
            for lin in net.linears:
                lin.bias.normal_(0, 0.1)
        params = [t for lin in net.linears for t in (lin.weight, lin.bias)]
        x = spec.sample(n, seed=4)
        with oracle_backend():
            op = FusedPDEResidual(spec, 100, 5, policy=policy, device="cpu", check_library=False)
            assert (op.world, op.rank) == (world, rank)
            op.set_points(x)
            beg, end = shard_bounds(n, rank, world)
            assert op.x.shape[0] == end - beg and op.n_global == n
            losses = op.losses(params)
            w = torch.tensor([0.5, 2.0, 1.5][:desc.n_pde])
            net.zero_grad()
            (losses * w).sum().backward(retain_graph=True)
            g1 = torch.cat([p.grad.reshape(-1) for p in params]).clone()
            net.zero_grad()
            losses[0].backward(retain_graph=True)          # a second backward with another seed (LRA pattern)
            g2 = torch.cat([p.grad.reshape(-1) for p in params]).clone()
            stats = dict(op.stats)
        q.put((rank, losses.detach().numpy(), g1.numpy(), g2.numpy(), stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,policy", [("Poisson2D_Classic", "auto"), ("NS2D_LidDriven", "auto"), ("NS2D_LidDriven", "eager")])
def test_two_rank_sharding_matches_single_process(name, policy):
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    from pinnacle_b200.nn import FNN
    n, world = 203, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, policy, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    # single-process reference on ALL rows
    torch.manual_seed(0)
    spec = problems.make(name)
    desc = spec.desc
    net = FNN([desc.d] + [100] * 5 + [desc.m])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    x = torch.as_tensor(spec.sample(n, seed=4))
    w = [0.5, 2.0, 1.5][:desc.n_pde]
    e0 = [1.0] + [0.0] * (desc.n_pde - 1)
    ol, og, _ = O.losses_and_grads(desc, flat, x, None, seeds=[w, e0])
    for rank, losses, g1, g2, stats in res:
        np.testing.assert_allclose(losses, ol.numpy(), rtol=2e-5)
        assert np.linalg.norm(g1 - og[0].numpy()) <= 2e-5 * np.linalg.norm(og[0].numpy())
        assert np.linalg.norm(g2 - og[1].numpy()) <= 2e-5 * np.linalg.norm(og[1].numpy())
        if policy == "eager" or desc.n_pde == 1:
            assert stats["fwd_bwd_calls"] == 1 and stats["allreduce_calls"] == 1     # one pass serves both backwards
        else:
            assert stats["fwd_calls"] == 1 and stats["fwd_bwd_calls"] == 2 and stats["allreduce_calls"] == 3
    # ranks end up bit-identical (same all-reduced buffer) => replicated optimizers stay in lock-step
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][1], res[1][1])


def test_shard_bounds_cover_all_rows():
    from pinnacle_b200.fused import shard_bounds
    for n in (1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            segs = [shard_bounds(n, r, world) for r in range(world)]
            assert segs[0][0] == 0 and segs[-1][1] == n
            assert all(segs[i][1] == segs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in segs]
            assert max(sizes) - min(sizes) <= 1


def _rows_worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pinnacle_b200 import fused, problems
        from pinnacle_b200.fused import shard_bounds
        from pinnacle_b200.nn import FNN
        from pinnacle_b200.solver import FusedPDESolver, ValueBC
        from util import oracle_backend
        torch.manual_seed(0)
        torch.set_num_threads(2)
        spec = problems.make("NS2D_LidDriven")
        net = FNN([2] + [100] * 5 + [3])
        with torch.no_grad():
            for lin in net.linears:
                lin.bias.normal_(0, 0.1)
        x = spec.sample(n, seed=4)
        xb = spec.sample(40, seed=5)
        terms = [ValueBC(0, 24, 0, np.ones((24, 1), np.float32)), ValueBC(24, 40, 2, np.zeros((16, 1), np.float32))]
        beg, end = shard_bounds(n, rank, world)
        orig = fused.FusedPDEResidual.__init__

        def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
            orig(self, spec, H, L, policy, "cpu", group, False)

        fused.FusedPDEResidual.__init__ = init_cpu
        with oracle_backend():
            s = FusedPDESolver(spec, net, x[beg:end], xb, terms, device="cpu", n_global=n).compile(torch.optim.Adam(net.parameters(), 1e-3))
            flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
            G, L = s.term_gradient_rows(flat, torch.eye(3))
            G, L = G.clone().numpy(), L.clone().numpy()
            # linear rows (NTK statistics, src/optimizer/ntk.py:24-58): gradient of the residual SUMS, sharded the same way
            Glin, _ = s.term_gradient_rows(flat, torch.ones(1, 3), linear=True)
            q.put((rank, G, L, dict(s.core.op.stats), Glin.clone().numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_rows_match_single_process():
    """The autograd-free row builder of the flat optimizers (FusedPDESolver.term_gradient_rows) under sharding: PDE rows are
    all-reduced once, boundary rows are replicated; every rank ends with the same global rows as a single process."""
    from pinnacle_b200 import fused, problems
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.solver import FusedPDESolver, ValueBC
    from util import oracle_backend
    n, world = 150, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    spec = problems.make("NS2D_LidDriven")
    net = FNN([2] + [100] * 5 + [3])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    x, xb = spec.sample(n, seed=4), spec.sample(40, seed=5)
    terms = [ValueBC(0, 24, 0, np.ones((24, 1), np.float32)), ValueBC(24, 40, 2, np.zeros((16, 1), np.float32))]
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            s = FusedPDESolver(spec, net, x, xb, terms, device="cpu").compile(torch.optim.Adam(net.parameters(), 1e-3))
            flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
            G, L = s.term_gradient_rows(flat, torch.eye(3))
            G, L = G.clone(), L.clone()
            Glin, _ = s.term_gradient_rows(flat, torch.ones(1, 3), linear=True)
    finally:
        fused.FusedPDEResidual.__init__ = orig
    for rank, Gr, Lr, stats, Glr in res:
        np.testing.assert_allclose(Lr, L.numpy(), rtol=2e-5)
        for t in range(5):
            assert np.linalg.norm(Gr[t] - G[t].numpy()) <= 2e-5 * np.linalg.norm(G[t].numpy()), t
        assert Glr.shape == (3, G.shape[1])                     # one PDE row (seed of ones) + two boundary rows
        for t in range(3):
            assert np.linalg.norm(Glr[t] - Glin[t].numpy()) <= 2e-5 * np.linalg.norm(Glin[t].numpy()), ("linear", t)
        assert stats["allreduce_calls"] == 2                    # one for the loss rows, one for the linear rows
    assert np.array_equal(res[0][4], res[1][4])                 # ranks agree bit for bit => replicated NTK weights stay in lock-step
    assert np.array_equal(res[0][1], res[1][1])


def _redraw_worker(rank, world, port, n, n_bc, n_dom, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pinnacle_b200 import fused, problems
        from pinnacle_b200.nn import FNN
        from pinnacle_b200.resample import Region
        from pinnacle_b200.solver import _StepCore
        from util import oracle_backend
        torch.manual_seed(0)
        torch.set_num_threads(2)
        spec = problems.make("PoissonBoltzmann2D")
        net = FNN([2] + [100] * 5 + [1])
        orig = fused.FusedPDEResidual.__init__

        def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
            orig(self, spec, H, L, policy, "cpu", group, False)

        fused.FusedPDEResidual.__init__ = init_cpu
        with oracle_backend():
            core = _StepCore(spec, net, "auto", "cpu", None, True)
            X = spec.sample(n_bc + n, seed=6).astype(np.float32).copy()
            core.stage(X, n_bc)
            region = Region(spec.bbox[0::2], spec.bbox[1::2], [(0, (0.0, 0.0), (0.4,))])
            core.redraw_domain_rows(X, n_bc, n_dom, region, key=21, draw=3)
            params = [t for lin in net.linears for t in (lin.weight, lin.bias)]
            losses = core.op.losses(params)
            q.put((rank, core.op.x.numpy().copy(), core.op.aux.numpy().copy(), losses.detach().numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_redraw_draws_the_rows_one_rank_would():
    """Sharded residual rows, device redraw (oracle standing in for the kernel): the two shards put together are the rows a
    single process draws -- a row is a function of its global index -- and the all-reduced loss is the single-process loss."""
    from oracle.sample_ref import sample_points
    from pinnacle_b200 import problems
    from pinnacle_b200.fused import shard_bounds
    n, n_bc, n_dom, world = 301, 17, 250, 2             # rank 0's shard holds kept rows AND redrawn rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_redraw_worker, args=(r, world, port, n, n_bc, n_dom, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = problems.make("PoissonBoltzmann2D")
    X = spec.sample(n_bc + n, seed=6).astype(np.float32)
    want = X[n_bc:].copy()
    want[n - n_dom:] = sample_points(n_dom, spec.bbox[0::2], spec.bbox[1::2], [(0, (0.0, 0.0), (0.4,))], key=21, draw=3)
    got = np.vstack([r[1] for r in res])
    assert np.array_equal(got, want)
    for r in res:
        beg, end = shard_bounds(n, r[0], world)
        np.testing.assert_allclose(r[2], spec.aux(torch.from_numpy(want[beg:end])).numpy(), rtol=1e-6)
    np.testing.assert_allclose(res[0][3], res[1][3], rtol=0, atol=0)
