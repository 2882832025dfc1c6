"""Row sharding over real GPUs with NCCL (skipped unless >= 2 CUDA devices): 2 ranks, each keeps its shard of the same
global point set; the all-reduced [grads | losses] must equal the single-GPU result and be bit-identical on both ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from pinnacle_b200 import problems
        from pinnacle_b200.fused import FusedPDEResidual, net_linear_params
        from pinnacle_b200.nn import FNN
        torch.manual_seed(0)
        spec = problems.make(name)
        net = FNN([spec.desc.d] + [100] * 5 + [spec.desc.m])
        with torch.no_grad():
            for lin in net.linears:
                lin.bias.normal_(0, 0.1)
        net = net.to(dev)
        params = net_linear_params(net)
        op = FusedPDEResidual(spec, device=dev).set_points(spec.sample(n, seed=4))
        losses = op.losses(params)
        losses.sum().backward()
        g = torch.cat([p.grad.reshape(-1) for p in params])
        q.put((rank, losses.detach().cpu().numpy(), g.cpu().numpy(), op.x.shape[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["Poisson2D_Classic", "NS2D_LidDriven"])
def test_two_gpu_sharding_matches_single_gpu(name):
    import torch.multiprocessing as mp
    from pinnacle_b200 import problems
    from pinnacle_b200.fused import FusedPDEResidual, net_linear_params
    from pinnacle_b200.nn import FNN
    n, world = 100003, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][3] + res[1][3] == n
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    spec = problems.make(name)
    net = FNN([spec.desc.d] + [100] * 5 + [spec.desc.m])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    net = net.to(dev)
    params = net_linear_params(net)
    op = FusedPDEResidual(spec, device=dev).set_points(spec.sample(n, seed=4))
    losses = op.losses(params)
    losses.sum().backward()
    g = torch.cat([p.grad.reshape(-1) for p in params]).cpu().numpy()
    for rank, l, gr, _ in res:
        np.testing.assert_allclose(l, losses.detach().cpu().numpy(), rtol=1e-5)
        assert np.linalg.norm(gr - g) <= 1e-5 * np.linalg.norm(g)
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][1], res[1][1])


def _opt_worker(rank, world, port, method, n, steps, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from pinnacle_b200.fused import shard_bounds
        beg, end = shard_bounds(n, rank, world)
        flat, w, losses = _run_flat_optimizer(method, n, steps, dev, beg, end)
        q.put((rank, flat, w, losses))
    finally:
        dist.destroy_process_group()


def _run_flat_optimizer(method, n, steps, dev, beg, end):
    """NS2D_LidDriven + FusedLRA / FusedNTK over FlatAdam on rows [beg, end) of an n-row problem (BASELINE configs[2])."""
    from pinnacle_b200 import problems
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.optim import FlatAdam, FusedLRA, FusedNTK
    from pinnacle_b200.solver import FusedPDESolver, ValueBC
    torch.manual_seed(0)
    spec = problems.make("NS2D_LidDriven")
    net = FNN([2] + [100] * 5 + [3])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    net = net.to(dev)
    x = spec.sample(n, seed=4)[beg:end]
    xb = spec.sample(512, seed=5)
    terms = [ValueBC(0, 256, 0, np.ones((256, 1), np.float32)), ValueBC(256, 512, 1, np.zeros((256, 1), np.float32))]
    w = np.ones(5)
    inner = FlatAdam(net.parameters(), 1e-3)
    opt = FusedLRA(inner, w, 3) if method == "lra" else FusedNTK(inner, w, 3)
    solver = FusedPDESolver(spec, net, x, xb, terms, device=dev, n_global=n).compile(opt, loss_weights=w)
    for _ in range(steps):
        solver.train_step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy()
    return flat, np.asarray(opt.sync_loss_weight(), dtype=np.float64).copy(), solver.opt.losses.detach().cpu().numpy()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("method", ["lra", "ntk"])
def test_two_gpu_flat_optimizers_stay_in_lock_step_and_match_single_gpu(method):
    """The loss-balancing optimizers see GLOBAL per-term gradients under sharding (loss rows and, for NTK, the linear rows
    of pinn_fwd_bwd_linear are all-reduced inside the step): both ranks end bit-identical and equal to one GPU on all rows."""
    import torch.multiprocessing as mp
    n, world, steps = 40003, 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_opt_worker, args=(r, world, port, method, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    flat, w, losses = _run_flat_optimizer(method, n, steps, torch.device("cuda:0"), 0, n)
    assert not np.allclose(w, 1.0)
    np.testing.assert_allclose(res[0][2], w, rtol=2e-3)
    np.testing.assert_allclose(res[0][3], losses, rtol=2e-3)
    assert np.linalg.norm(res[0][1] - flat) <= 1e-4 * np.linalg.norm(flat)
