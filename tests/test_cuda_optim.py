"""Optimizer-side kernels over the flat buffers on the GPU: FlatAdam against torch.optim.Adam,
FusedLRA against a line-by-line restatement of the reference's LR_Adaptor driven through autograd, the statistics kernel
against torch reductions, and CUDA-graph replay of the autograd-free step."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(desc, dev, seed=0):
    from pinnacle_b200.nn import FNN
    torch.manual_seed(seed)
    net = FNN([desc.d] + [100] * 5 + [desc.m])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    return net.to(dev)


def _flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


def _ns_problem(dev, n=4096, nb=256, seed=0):
    """NS2D_LidDriven residual rows + four value-type boundary terms (u, v on two walls), BASELINE configs[2] at test size."""
    from pinnacle_b200 import problems
    from pinnacle_b200.solver import ValueBC
    spec = problems.make("NS2D_LidDriven")
    x = spec.sample(n, seed=seed + 1)
    rs = np.random.RandomState(seed + 2)
    xb = rs.uniform(0, 1, (4 * nb, 2)).astype(np.float32)
    xb[:nb, 1] = 1.0
    xb[nb:2 * nb, 1] = 0.0
    xb[2 * nb:3 * nb, 0] = 0.0
    xb[3 * nb:, 0] = 1.0
    lid = (4.0 * xb[:nb, :1] * (1.0 - xb[:nb, :1])).astype(np.float32)
    terms = [ValueBC(0, nb, 0, lid), ValueBC(0, nb, 1, np.zeros((nb, 1), np.float32)),
             ValueBC(nb, 4 * nb, 0, np.zeros((3 * nb, 1), np.float32)), ValueBC(nb, 4 * nb, 1, np.zeros((3 * nb, 1), np.float32)),
             ValueBC(4 * nb - 1, 4 * nb, 2, np.zeros((1, 1), np.float32))]       # one-point pressure pin (ns.py PointSetBC)
    return spec, x, xb, terms


def test_grad_stats_matches_torch():
    from pinnacle_b200 import capi
    dev = torch.device("cuda:0")
    g = torch.randn(41003, device=dev) * 3.0
    out = capi.grad_stats(g).cpu().double().numpy()
    ref = np.array([g.abs().max().item(), g.abs().double().sum().item(), (g.double() ** 2).sum().item()])
    np.testing.assert_allclose(out, ref, rtol=2e-6)


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_follows_torch_adam(wd):
    """Same gradients in, same parameters out: 25 updates on the fused Poisson loss, generic closure protocol."""
    from pinnacle_b200 import problems
    from pinnacle_b200.optim import FlatAdam
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec = problems.make("Poisson2D_Classic")
    x = spec.sample(2048, seed=3)
    net_a, net_b = _net(spec.desc, dev, 5), _net(spec.desc, dev, 5)
    sa = FusedPDESolver(spec, net_a, x).compile(FlatAdam(net_a.parameters(), lr=1e-3, weight_decay=wd))
    sb = FusedPDESolver(spec, net_b, x).compile(torch.optim.Adam(net_b.parameters(), lr=1e-3, weight_decay=wd))
    assert sa.opt.flat.data_ptr() == next(net_a.parameters()).data_ptr()
    for _ in range(25):
        sa.train_step()
        sb.train_step()
    la, lb = sa.opt.losses.detach().cpu().numpy(), sb.opt.losses.detach().cpu().numpy()
    np.testing.assert_allclose(la, lb, rtol=2e-4)
    pa, pb = _flat(net_a), _flat(net_b)
    assert float((pa - pb).norm() / pb.norm()) < 2e-5
    assert float(sa.opt.step_dev[0]) == 25.0


def test_autograd_free_step_matches_autograd_step_and_replays_as_graph():
    from pinnacle_b200.optim import FlatAdam
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec, x, xb, terms = _ns_problem(dev)
    w = np.array([1.0, 1.0, 1.0, 2.0, 3.0, 0.5, 1.5, 10.0])
    nets = [_net(spec.desc, dev, 7) for _ in range(3)]
    sa = FusedPDESolver(spec, nets[0], x, xb, terms).compile(FlatAdam(nets[0].parameters(), lr=1e-3), loss_weights=w)
    sb = FusedPDESolver(spec, nets[1], x, xb, terms).compile(torch.optim.Adam(nets[1].parameters(), lr=1e-3), loss_weights=w.copy())
    sc = FusedPDESolver(spec, nets[2], x, xb, terms).compile(FlatAdam(nets[2].parameters(), lr=1e-3), loss_weights=w.copy())
    sc.capture_step()            # 3 warm-up steps inside, then every train_step() is a graph replay
    for _ in range(3):
        sa.train_step()
        sb.train_step()
    for k in range(6):
        sa.train_step()
        sb.train_step()
        sc.train_step()
        la, lb = sa.opt.losses.detach().cpu().numpy(), sb.opt.losses.detach().cpu().numpy()
        np.testing.assert_allclose(la, lb, rtol=5e-4)
    # the captured solver has made the same number of updates (capture itself executes nothing): 3 warm-up + 6 replays
    assert float(sc.opt.step_dev[0]) == float(sa.opt.step_dev[0]) == 9.0
    np.testing.assert_allclose(sc.opt.losses.detach().cpu().numpy(), sa.opt.losses.detach().cpu().numpy(), rtol=5e-4)
    pa, pb, pc = _flat(nets[0]), _flat(nets[1]), _flat(nets[2])
    assert float((pa - pc).norm() / pa.norm()) < 1e-4
    assert sa.core.op.stats["fwd_calls"] == 0         # no autograd forward: one fwd_bwd per term per step


class _LRAdaptorPort(torch.optim.Optimizer):
    """Test-side restatement of the reference's LR_Adaptor.step (src/optimizer/lr_adaptor.py:27-64) on torch tensors
    (device-agnostic: the reference's `torch.as_tensor(self.loss_weight)` of a host array becomes an explicit device copy)."""

    def __init__(self, optimizer, loss_weight, num_pde, alpha=0.1):
        super().__init__(optimizer.param_groups, dict(alpha=alpha))
        self.optimizer, self.loss_weight, self.num_pde, self.alpha = optimizer, loss_weight, num_pde, alpha

    @torch.no_grad()
    def step(self, closure):
        with torch.enable_grad():
            closure(skip_backward=True)
            dev = self.losses.device
            losses = self.losses / torch.as_tensor(self.loss_weight, dtype=torch.float32).to(dev)
            pde_loss = torch.sum(losses[:self.num_pde])
        self.zero_grad()
        pde_loss.backward(retain_graph=True)
        m_grad_r = max(torch.max(torch.abs(p.grad)).item() for g in self.param_groups for p in g["params"] if p.grad is not None)
        for i in range(self.num_pde, len(self.loss_weight)):
            with torch.enable_grad():
                self.zero_grad()
                losses[i].backward(retain_graph=True)
            s, c = 0.0, 0
            for g in self.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        s += torch.sum(torch.abs(p.grad)).item()
                        c += p.grad.numel()
            lambda_hat = m_grad_r / ((s / c) * self.loss_weight[i])
            self.loss_weight[i] = (1 - self.alpha) * self.loss_weight[i] + self.alpha * lambda_hat
        with torch.enable_grad():
            self.zero_grad()
            total = torch.sum(losses * torch.as_tensor(self.loss_weight, dtype=torch.float32).to(dev))
            total.backward()
        self.optimizer.step()
        return total


@pytest.mark.parametrize("pde_w", [1.0, None])
def test_fused_lra_follows_reference_algorithm(pde_w):
    """NS2D_LidDriven + learning-rate annealing (BASELINE configs[2]): the row-based FusedLRA against the reference's
    algorithm run through autograd (one backward per term, host-side statistics) on the same fused losses."""
    from pinnacle_b200.optim import FlatAdam, FusedLRA
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec, x, xb, terms = _ns_problem(dev)
    w0 = np.ones(8) if pde_w is not None else np.array([1.0, 2.0, 0.5, 1.0, 1.0, 1.0, 1.0, 1.0])
    net_a, net_b = _net(spec.desc, dev, 9), _net(spec.desc, dev, 9)
    wa, wb = w0.copy(), w0.copy()
    sa = FusedPDESolver(spec, net_a, x, xb, terms)
    sa.compile(FusedLRA(FlatAdam(net_a.parameters(), lr=1e-3), wa, 3), loss_weights=wa)
    sb = FusedPDESolver(spec, net_b, x, xb, terms)
    sb.compile(_LRAdaptorPort(torch.optim.Adam(net_b.parameters(), lr=1e-3), wb, 3), loss_weights=wb)
    for k in range(6):
        ta = sa.train_step()
        tb = sb.train_step()
        assert abs(float(ta) - float(tb)) <= 2e-3 * abs(float(tb)), (k, float(ta), float(tb))
        np.testing.assert_allclose(sa.opt.sync_loss_weight(), wb, rtol=2e-3)
    assert not np.allclose(wb[3:], 1.0)
    np.testing.assert_array_equal(wa[:3], w0[:3])
    pa, pb = _flat(net_a), _flat(net_b)
    assert float((pa - pb).norm() / pb.norm()) < 2e-3


class _MultiAdamPort(torch.optim.Optimizer):
    """Test-side restatement of the reference's MultiAdam.step + sadam (src/optimizer/multiadam.py:54-122, 216-330) with its
    benchmark.py:108 configuration (amsgrad off, no aggregated momentum, no parameter scheduler), flat over all parameters."""

    def __init__(self, params, lr=1e-3, betas=(0.99, 0.99), eps=1e-8, loss_group_idx=None, group_weights=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.loss_group_idx = list(loss_group_idx or [])
        self.n_groups = len(self.loss_group_idx) + 1
        self.gw = [1.0 / self.n_groups] * self.n_groups if group_weights is None else list(group_weights)
        self.m = [[torch.zeros_like(p) for p in params] for _ in range(self.n_groups)]
        self.v = [[torch.zeros_like(p) for p in params] for _ in range(self.n_groups)]
        self.t = 0

    @torch.no_grad()
    def step(self, closure):
        with torch.enable_grad():
            closure(skip_backward=True)
            losses = self.losses
            idx = [0] + self.loss_group_idx + [len(losses)]
            grouped = [torch.sum(losses[idx[i]:idx[i + 1]]) for i in range(len(idx) - 1)]
        params = self.param_groups[0]["params"]
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        grads = []
        for loss in grouped:
            self.zero_grad()
            with torch.enable_grad():
                loss.backward(retain_graph=True)
            grads.append([torch.zeros_like(p) if p.grad is None else p.grad.clone() for p in params])
        self.t += 1
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for j, p in enumerate(params):
            upd = torch.zeros_like(p)
            for k in range(self.n_groups):
                self.m[k][j].mul_(b1).add_(grads[k][j], alpha=1 - b1)
                self.v[k][j].mul_(b2).addcmul_(grads[k][j], grads[k][j], value=1 - b2)
                upd += self.gw[k] * self.m[k][j] / ((self.v[k][j].sqrt() / bc2 ** 0.5) + g["eps"])
            p -= (g["lr"] / bc1) * upd
        return sum(grouped)


@pytest.mark.parametrize("groups", [[3], [1, 3]])
def test_flat_multiadam_follows_reference_algorithm(groups):
    """benchmark.py --method multiadam (loss_group_idx = [num_pde], betas (0.99, 0.99)): the one-launch FlatMultiAdam on the
    gradient rows against the reference algorithm run through autograd; also with a group boundary inside the PDE terms."""
    from pinnacle_b200.optim import FlatMultiAdam
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec, x, xb, terms = _ns_problem(dev)
    w = np.array([1.0, 2.0, 0.5, 1.0, 3.0, 1.0, 0.25, 10.0])
    net_a, net_b = _net(spec.desc, dev, 21), _net(spec.desc, dev, 21)
    sa = FusedPDESolver(spec, net_a, x, xb, terms).compile(
        FlatMultiAdam(net_a.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=groups), loss_weights=w)
    sb = FusedPDESolver(spec, net_b, x, xb, terms).compile(
        _MultiAdamPort(net_b.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=groups), loss_weights=w.copy())
    for k in range(8):
        sa.train_step()
        sb.train_step()
        np.testing.assert_allclose(sa.opt.losses.detach().cpu().numpy(), sb.opt.losses.detach().cpu().numpy(), rtol=1e-3)
    pa, pb = _flat(net_a), _flat(net_b)
    assert float((pa - pb).norm() / pb.norm()) < 1e-4
    assert float(sa.opt.step_dev[0]) == 8.0
    # graph replay of the same step
    net_c = _net(spec.desc, dev, 21)
    sc = FusedPDESolver(spec, net_c, x, xb, terms).compile(
        FlatMultiAdam(net_c.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=groups), loss_weights=w.copy())
    sc.capture_step(warmup=3)
    for _ in range(5):
        sc.train_step()
    assert float((_flat(net_c) - pa).norm() / pa.norm()) < 1e-4


class _NTKPort(torch.optim.Optimizer):
    """Test-side restatement of LR_Adaptor_NTK.step (src/optimizer/ntk.py:22-66): `sums_fn()` returns the differentiable sums of
    all PDE residuals and of all boundary errors (data.losses with loss_fn = (y - x).sum())."""

    def __init__(self, optimizer, loss_weight, num_pde, sums_fn):
        super().__init__(optimizer.param_groups, {})
        self.optimizer, self.loss_weight, self.num_pde, self.sums_fn = optimizer, loss_weight, num_pde, sums_fn

    @torch.no_grad()
    def step(self, closure):
        with torch.enable_grad():
            loss_r, loss_b = self.sums_fn()
        sq = []
        for loss in (loss_r, loss_b):
            self.zero_grad()
            with torch.enable_grad():
                loss.backward(retain_graph=True)
            sq.append(sum(float((p.grad ** 2).sum()) for g in self.param_groups for p in g["params"] if p.grad is not None))
        m_r, m_b = sq
        for i in range(len(self.loss_weight)):
            self.loss_weight[i] = (m_r + m_b) / (m_r if i < self.num_pde else m_b)
        return self.optimizer.step(closure)          # the solver's closure reports its losses to `self` (solver.opt)


def test_fused_ntk_follows_reference_algorithm():
    """benchmark.py --method ntk: FusedNTK (linear gradient rows from pinn_fwd_bwd_linear, weights on the device) against the
    reference algorithm with stock autograd statistics (oracle residuals on the GPU + net forward on the BC rows)."""
    from oracle import autograd_ref as O
    from pinnacle_b200.optim import FlatAdam, FusedNTK
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec, x, xb, terms = _ns_problem(dev, n=1024, nb=64)
    net_a, net_b = _net(spec.desc, dev, 23), _net(spec.desc, dev, 23)
    wa, wb = np.ones(8), np.ones(8)
    xt, xbt = torch.as_tensor(x).to(dev), torch.as_tensor(xb).to(dev)

    def sums():
        flat = torch.cat([p.reshape(-1) for p in net_b.parameters()])
        _, res, _ = O.pde_losses(spec.desc, flat, xt, None)
        yb = net_b(xbt)
        lb = sum((yb[t.beg:t.end, t.component:t.component + 1] - torch.as_tensor(t.values, device=dev)).sum() for t in terms)
        return res.sum(), lb

    sa = FusedPDESolver(spec, net_a, x, xb, terms).compile(FusedNTK(FlatAdam(net_a.parameters(), lr=1e-3), wa, 3), loss_weights=wa)
    inner_b = torch.optim.Adam(net_b.parameters(), lr=1e-3)
    sb = FusedPDESolver(spec, net_b, x, xb, terms).compile(_NTKPort(inner_b, wb, 3, sums), loss_weights=wb)
    for k in range(4):
        sa.train_step()
        sb.train_step()
        np.testing.assert_allclose(sa.opt.sync_loss_weight(), wb, rtol=2e-3)
        np.testing.assert_allclose(sa.opt.losses.detach().cpu().numpy(), sb.opt.losses.detach().cpu().numpy(), rtol=2e-3)
    assert wb[0] != 1.0 and wb[3] != 1.0
    assert float((_flat(net_a) - _flat(net_b)).norm() / _flat(net_b).norm()) < 1e-3


@pytest.mark.parametrize("kind", ["adam", "lra", "multiadam"])
def test_flat_optimizer_checkpoint_round_trip(kind, tmp_path):
    """model.save / restore (deepxde/model.py:933-987) store net.state_dict() and opt.state_dict(): a run restored from a
    checkpoint continues exactly like the uninterrupted one (moments, device step counter, device loss weights included)."""
    from pinnacle_b200.optim import FlatAdam, FlatMultiAdam, FusedLRA
    from pinnacle_b200.solver import FusedPDESolver
    dev = torch.device("cuda:0")
    spec, x, xb, terms = _ns_problem(dev, n=1024, nb=64)

    def make(seed):
        net = _net(spec.desc, dev, seed)
        w = np.ones(8)
        if kind == "adam":
            opt = FlatAdam(net.parameters(), lr=1e-3)
        elif kind == "lra":
            opt = FusedLRA(FlatAdam(net.parameters(), lr=1e-3), w, 3)
        else:
            opt = FlatMultiAdam(net.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=[3])
        return net, FusedPDESolver(spec, net, x, xb, terms).compile(opt, loss_weights=w)

    net_a, sa = make(31)
    for _ in range(4):
        sa.train_step()
    path = str(tmp_path / "ckpt.pt")
    torch.save({"model_state_dict": net_a.state_dict(), "optimizer_state_dict": sa.opt.state_dict()}, path)
    for _ in range(3):
        sa.train_step()
    net_b, sb = make(99)                         # different initial weights: everything must come from the checkpoint
    ck = torch.load(path, weights_only=False)
    net_b.load_state_dict(ck["model_state_dict"])
    sb.opt.load_state_dict(ck["optimizer_state_dict"])
    assert next(net_b.parameters()).data_ptr() == (sb.opt.optimizer if kind == "lra" else sb.opt).flat.data_ptr()   # still views of the flat buffer
    for _ in range(3):
        sb.train_step()
    np.testing.assert_allclose(sb.opt.losses.detach().cpu().numpy(), sa.opt.losses.detach().cpu().numpy(), rtol=1e-6)
    assert float((_flat(net_a) - _flat(net_b)).norm() / _flat(net_a).norm()) < 1e-7


def test_flat_adam_checkpoint_interchanges_with_torch_adam():
    """A checkpoint written by the reference's torch.optim.Adam (benchmark.py:107) resumes under FlatAdam with its moments and
    step count, and a FlatAdam checkpoint resumes under torch.optim.Adam: in both directions the continued run matches the
    uninterrupted stock-Adam run (deepxde/model.py:945-948, 979-987 store / load opt.state_dict())."""
    from pinnacle_b200 import capi
    from pinnacle_b200.optim import FlatAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    tgt = torch.randn(64, 3, device=dev)
    xin = torch.randn(64, 7, device=dev)

    def net():
        torch.manual_seed(6)
        return torch.nn.Sequential(torch.nn.Linear(7, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3)).to(dev)

    def run(n, opt, steps):
        for _ in range(steps):
            opt.zero_grad()
            ((n(xin) - tgt) ** 2).mean().backward()
            opt.step()

    flat = lambda n: torch.cat([p.detach().reshape(-1) for p in n.parameters()])
    ref = net(); oref = torch.optim.Adam(ref.parameters(), lr=3e-3, weight_decay=1e-4)
    run(ref, oref, 5)
    sd_torch = copy.deepcopy(oref.state_dict())          # state_dict() hands out the live state tensors
    w5 = [p.detach().clone() for p in ref.parameters()]
    run(ref, oref, 4)
    # torch.optim.Adam -> FlatAdam
    a = net()
    with torch.no_grad():
        for p, w in zip(a.parameters(), w5):
            p.copy_(w)
    oa = FlatAdam(a.parameters(), lr=1.0)                 # hyper-parameters come from the checkpoint's param group
    oa.load_state_dict(sd_torch)
    assert float(oa.step_dev[0]) == 5.0 and float(oa.exp_avg.abs().sum()) > 0
    run(a, oa, 4)
    assert float((flat(a) - flat(ref)).norm() / flat(ref).norm()) < 1e-6
    # FlatAdam -> torch.optim.Adam
    b = net(); ob = FlatAdam(b.parameters(), lr=3e-3, weight_decay=1e-4)
    run(b, ob, 5)
    sd_flat = copy.deepcopy(ob.state_dict())
    c = net()
    with torch.no_grad():
        for p, q in zip(c.parameters(), b.parameters()):
            p.copy_(q)
    oc = torch.optim.Adam(c.parameters(), lr=1.0)
    oc.load_state_dict(sd_flat)
    run(c, oc, 4)
    assert float((flat(c) - flat(ref)).norm() / flat(ref).norm()) < 1e-6
    # optimizer state in neither form is an error, not a silent reset
    bad = dict(sd_torch); bad["state"] = {k: {"momentum_buffer": v["exp_avg"]} for k, v in sd_torch["state"].items()}
    with pytest.raises(capi.FusedOpError):
        FlatAdam(net().parameters(), lr=1e-3).load_state_dict(bad)


def test_workspace_is_per_stream_and_survives_graph_capture():
    """the scratch buffer a captured step baked into its CUDA graph must never be handed to the allocator again,
    and two streams must not share stash / partial accumulators."""
    from pinnacle_b200 import capi, problems
    dev = torch.device("cuda:0")
    spec = problems.make("Poisson2D_Classic")
    desc = spec.desc.with_net(100, 5)
    net = _net(spec.desc, dev, 3)
    flat = _flat(net).contiguous()
    x = torch.as_tensor(spec.sample(3000, seed=2)).to(dev)
    seeds = torch.ones(1, 1, device=dev)
    ref = capi.fwd_bwd(desc, flat, x, None, 1.0 / 3000, seeds).clone()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        capi.fwd_bwd(desc, flat, x, None, 1.0 / 3000, seeds)
    torch.cuda.synchronize()
    out = torch.empty_like(ref)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        capi.fwd_bwd(desc, flat, x, None, 1.0 / 3000, seeds, out=out)
    n_captured = len(capi._WS._captured)
    assert n_captured >= 1
    # a later, larger request on the default stream (more seeds => bigger workspace) must not disturb the captured buffer
    big = capi.fwd_bwd(desc, flat, x, None, 1.0 / 3000, torch.ones(3, 1, device=dev))
    junk = [torch.full((1 << 22,), float("nan"), device=dev) for _ in range(8)]      # would land in a freed workspace
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref) and torch.equal(big[:desc.n_params], ref[:desc.n_params])
    # concurrent calls on two streams
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        o1 = [capi.fwd_bwd(desc, flat, x, None, 1.0 / 3000, seeds) for _ in range(4)]
    with torch.cuda.stream(s2):
        o2 = [capi.fwd_bwd(desc, flat, x, None, 1.0 / 3000, seeds) for _ in range(4)]
    torch.cuda.synchronize()
    assert all(torch.equal(o, ref) for o in o1 + o2)
    del junk
