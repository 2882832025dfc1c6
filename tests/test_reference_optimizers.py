"""The reference's own optimizer classes (src/optimizer/{lr_adaptor,multiadam,adam_lbfgs,ntk}.py, unmodified) driving a
patched dde.Model, against the same optimizer driving the unpatched model: same weights, same points, a few steps.
BASELINE.json configs[2] (NS2D_LidDriven + LRA) and configs[3] (Kuramoto-Sivashinsky, Adam then L-BFGS) at test size.

CPU, reference tree mounted only (the oracle stands in for the C ABI as the compute backend of the host logic)."""
import copy

import numpy as np
import pytest
import torch

import refshim
from util import oracle_backend

needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


def _build(name, method, seed=0):
    dde, _ = refshim.load()
    from src.optimizer import MultiAdam, LR_Adaptor, LR_Adaptor_NTK, Adam_LBFGS   # the reference's classes
    torch.manual_seed(seed)
    dde.config.set_random_seed(seed)
    pde = refshim.construct(name)
    pde.training_points(domain=128, boundary=48, **({"initial": 48} if hasattr(pde, "num_initial_points") else {}))
    pde.num_test_points = 64
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)       # zero biases make a BC gradient mean of exactly 0 -> inf in LRA
    loss_weights = np.ones(pde.num_loss)
    opt = torch.optim.Adam(net.parameters(), 1e-3)                                      # benchmark.py:107-115
    if method == "multiadam":
        opt = MultiAdam(net.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=[pde.num_pde])
    elif method == "lra":
        opt = LR_Adaptor(opt, loss_weights, pde.num_pde)
    elif method == "ntk":
        opt = LR_Adaptor_NTK(opt, loss_weights, pde)
    elif method == "lbfgs":
        opt = Adam_LBFGS(net.parameters(), switch_epoch=3, adam_param={"lr": 1e-3})
    model = pde.create_model(net)
    model.compile(opt, loss_weights=loss_weights)
    net.auxiliary_vars = None
    return model, net, loss_weights


def _run(model, steps):
    hist = []
    x = model.data.train_x
    for _ in range(steps):
        model._train_step(x, None, None)
        hist.append(model.opt.losses.detach().cpu().numpy().astype(np.float64).copy())
    return np.array(hist)


@needs_ref
@pytest.mark.parametrize("name,method,steps", [
    ("NS2D_LidDriven", "lra", 4),                       # configs[2]
    ("KuramotoSivashinskyEquation", "lbfgs", 4),        # configs[3]: 2 Adam steps, then 2 L-BFGS steps (<= 20 closure calls each)
    ("Burgers1D", "multiadam", 4),
    ("Poisson2D_Classic", "ntk", 3),
    ("Heat2D_Multiscale", "lra", 3),
])
def test_reference_optimizer_on_patched_model_follows_unpatched(name, method, steps):
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    ref_model, ref_net, ref_w = _build(name, method)
    new_model, new_net, new_w = _build(name, method)
    for a, b in zip(ref_net.parameters(), new_net.parameters()):
        assert torch.equal(a, b)
    assert np.array_equal(ref_model.data.train_x, new_model.data.train_x)
    ref_hist = _run(ref_model, steps)

    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(new_model)
            new_hist = _run(new_model, steps)
    finally:
        fused.FusedPDEResidual.__init__ = orig

    assert np.all(np.isfinite(ref_hist)) and np.all(np.isfinite(new_hist)), (ref_hist, new_hist)
    # step 1 sees identical weights: per-term (weighted) losses to parity tolerance; later steps compound the optimizer's
    # sensitivity to 1e-7-level gradient differences (L-BFGS line search, LRA's max|g| / mean|g| ratio)
    np.testing.assert_allclose(new_hist[0], ref_hist[0], rtol=5e-5)
    np.testing.assert_allclose(new_hist, ref_hist, rtol=2e-2 if method == "lbfgs" else 5e-3)
    if method in ("lra", "ntk"):
        # the adapted loss weights (mutated in place by the optimizer, read live by outputs_losses) follow too
        np.testing.assert_allclose(np.asarray(new_w, dtype=np.float64), np.asarray(ref_w, dtype=np.float64), rtol=5e-3)
        assert not np.allclose(np.asarray(ref_w, dtype=np.float64), 1.0)
    for a, b in zip(ref_net.parameters(), new_net.parameters()):
        assert (a - b).norm() <= (2e-2 if method == "lbfgs" else 2e-3) * a.norm().clamp_min(1e-6)


@needs_ref
@pytest.mark.parametrize("name", ["Burgers1D", "NS2D_LidDriven", "Poisson2D_ManyArea"])
def test_predict_operator_pde_and_rar_wrapper_on_patched_model(name):
    """model.predict(X, operator=pde.pde) -- the residual query of RAR (src/utils/rar.py:20-21) -- is served by the fused
    forward-only path with the reference's return structure, and the reference's rar_wrapper picks the same anchor."""
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    model, net, _ = _build(name, "adam")       # loads the reference (sys.path) first
    from src.utils.rar import rar_wrapper
    pde = model.pde
    X = model.data.train_x
    ref = model.predict(X, operator=pde.pde)
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            new = model.predict(X, operator=pde.pde)
            assert isinstance(new, list) == isinstance(ref, list)
            r, n = np.asarray(ref, dtype=np.float64), np.asarray(new, dtype=np.float64)
            assert r.shape == n.shape
            assert np.linalg.norm(r - n) <= 2e-5 * np.linalg.norm(r)
            # plain prediction and foreign operators keep the reference's code path
            np.testing.assert_array_equal(model.predict(X[:5]), net(torch.as_tensor(X[:5])).detach().numpy())
            other = model.predict(X[:5], operator=lambda x, y: y * 2.0)
            np.testing.assert_allclose(other, 2.0 * model.predict(X[:5]), rtol=1e-6)
            # the reference's RAR loop on the patched model: train, query residuals, add the worst point as an anchor, train
            n_before = model.data.train_x_all.shape[0]
            model.train = rar_wrapper(pde, model, {"interval": 2, "count": 1})
            model.train(iterations=4, display_every=2)
            assert model.data.train_x_all.shape[0] == n_before + 1
            assert np.all(np.isfinite(model.opt.losses.detach().numpy()))
    finally:
        fused.FusedPDEResidual.__init__ = orig


@needs_ref
@pytest.mark.parametrize("name,kind", [("Burgers1D", "laaf"), ("Poisson2D_Classic", "gaaf"), ("NS2D_LidDriven", "laaf")])
def test_adaptive_activation_nets_on_patched_model(name, kind):
    """benchmark.py --method laaf / gaaf: the reference's DNN_LAAF / DNN_GAAF (src/model/laaf.py, unmodified) under a patched
    model -- the fused op runs on the equivalent FNN (W' = diag(10 a) W, b' = 10 a b) and autograd carries the gradient
    back to W, b AND the slopes a.  Losses, every parameter gradient, the _test() path and three Adam steps."""
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    dde, _ = refshim.load()
    from src.model.laaf import DNN_GAAF, DNN_LAAF

    def build():
        torch.manual_seed(3)
        dde.config.set_random_seed(3)
        pde = refshim.construct(name)
        pde.training_points(domain=128, boundary=48, **({"initial": 48} if hasattr(pde, "num_initial_points") else {}))
        pde.num_test_points = 64
        net = (DNN_LAAF if kind == "laaf" else DNN_GAAF)(5, 100, pde.input_dim, pde.output_dim).float()
        with torch.no_grad():
            net.a.mul_(0.3)            # slopes 10*a of order one (the xavier init gives |10 a| up to ~3: saturated tanh everywhere)
        model = pde.create_model(net)
        model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
        net.auxiliary_vars = None
        return model, net

    ref_model, ref_net = build()
    new_model, new_net = build()
    x = ref_model.data.train_x
    _, ref_losses = ref_model.outputs_losses_train(x, None)
    ref_net.zero_grad()
    ref_losses.sum().backward()
    ref_test = ref_model._outputs_losses(False, ref_model.data.test_x, None, None)
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(new_model)
            assert new_model.fused_core.dynamic_params
            _, new_losses = new_model.outputs_losses_train(x, None)
            new_net.zero_grad()
            new_losses.sum().backward()
            np.testing.assert_allclose(new_losses.detach().numpy(), ref_losses.detach().numpy(), rtol=5e-5)
            names = [n for n, _ in ref_net.named_parameters()]
            assert "a" in names
            for (n, p), (_, q) in zip(ref_net.named_parameters(), new_net.named_parameters()):
                rg = torch.zeros_like(p) if p.grad is None else p.grad
                ng = torch.zeros_like(q) if q.grad is None else q.grad
                assert (ng - rg).norm() <= 5e-5 * max(float(rg.norm()), 1e-12), (n, float((ng - rg).norm()), float(rg.norm()))
            new_test = new_model._outputs_losses(False, new_model.data.test_x, None, None)
            np.testing.assert_allclose(new_test[1], ref_test[1], rtol=5e-5)
            np.testing.assert_allclose(new_test[0], ref_test[0], rtol=1e-5, atol=1e-6)
            ref_hist, new_hist = _run(ref_model, 3), _run(new_model, 3)
            np.testing.assert_allclose(new_hist, ref_hist, rtol=5e-3)
    finally:
        fused.FusedPDEResidual.__init__ = orig


@needs_ref
def test_swap_optimizer_reads_the_reference_classes_correctly(monkeypatch):
    """enable_fused(fuse_optimizer=True) maps the reference's own optimizer objects (benchmark.py:107-115) to their flat-buffer
    equivalents with the same hyper-parameters.  The flat optimizers need CUDA, so recorders stand in for them here; the
    GPU side of the swap is tests/test_cuda_solver.py::test_enable_fused_swaps_in_the_flat_optimizers."""
    from pinnacle_b200 import optim, solver
    dde, _ = refshim.load()
    from src.optimizer import MultiAdam, LR_Adaptor, LR_Adaptor_NTK, Adam_LBFGS
    made = []

    class Rec:
        def __init__(self, *a, **k):
            made.append((type(self).__name__, a, k))

    for nm in ("FlatAdam", "FlatMultiAdam", "FusedLRA", "FusedNTK"):
        monkeypatch.setattr(optim, nm, type(nm, (Rec,), {}))
    net = dde.nn.FNN([2] + [100] * 5 + [1], "tanh", "Glorot normal").float()
    w = np.ones(4)
    adam = torch.optim.Adam(net.parameters(), 3e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=1e-4)
    out = solver._swap_optimizer(adam, net, w, 1)
    assert type(out).__name__ == "FlatAdam" and made[-1][2] == dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=1e-4)
    out = solver._swap_optimizer(LR_Adaptor(torch.optim.Adam(net.parameters(), 1e-3), w, 1, alpha=0.3), net, w, 1)
    assert type(out).__name__ == "FusedLRA"
    assert made[-1][1][1] is w and made[-1][1][2] == 1 and made[-1][2] == dict(alpha=0.3)
    out = solver._swap_optimizer(MultiAdam(net.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=[1]), net, w, 1)
    assert type(out).__name__ == "FlatMultiAdam"
    assert made[-1][2]["loss_group_idx"] == [1] and made[-1][2]["betas"] == (0.99, 0.99) and made[-1][2]["group_weights"] == [0.5, 0.5]
    pde = refshim.construct("Poisson2D_Classic")
    pde.training_points(domain=32, boundary=16)
    pde.create_model(net)
    out = solver._swap_optimizer(LR_Adaptor_NTK(torch.optim.Adam(net.parameters(), 1e-3), w, pde), net, w, 1)
    assert type(out).__name__ == "FusedNTK" and made[-1][1][1] is w and made[-1][1][2] == 1
    # no flat equivalent: left alone
    assert solver._swap_optimizer(Adam_LBFGS(net.parameters(), switch_epoch=5), net, w, 1) is None
    assert solver._swap_optimizer(torch.optim.Adam(net.parameters(), 1e-3, amsgrad=True), net, w, 1) is None
    assert solver._swap_optimizer(MultiAdam(net.parameters(), loss_group_idx=[1], agg_momentum=True, agg_betas=(0.9, 0.9)), net, w, 1) is None
    assert solver._swap_optimizer(LR_Adaptor(torch.optim.Adam(net.parameters(), 1e-3), w, 2), net, w, 1) is None      # num_pde mismatch


@needs_ref
@pytest.mark.parametrize("name", ["PoissonInv", "HeatInv"])
def test_inverse_problems_with_their_recommended_pfnn(name):
    """The inverse problems with the net their class recommends (src/pde/inverse.py:30-32,143-148: a parallel FNN of two
    50-wide sub-networks, HeatInv's with an input mask): the fused op runs on the equivalent dense 100-wide FNN (block-diagonal
    weights), autograd returns the gradient of the blocks to the branch parameters."""
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    dde, _ = refshim.load()

    def build():
        torch.manual_seed(5)
        np.random.seed(5)
        dde.config.set_random_seed(5)
        pde = refshim.construct(name)
        pde.training_points(domain=96, boundary=32, **({"initial": 32} if hasattr(pde, "num_initial_points") else {}))
        pde.num_test_points = 48
        net = pde.recommend_net.float()
        with torch.no_grad():
            for lay in net.layers:
                for f in lay:
                    f.bias.normal_(0, 0.1)
        model = pde.create_model(net)
        model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
        net.auxiliary_vars = None
        return model, net

    ref_model, ref_net = build()
    new_model, new_net = build()
    for a, b in zip(ref_net.parameters(), new_net.parameters()):
        assert torch.equal(a, b)
    x = ref_model.data.train_x
    assert np.array_equal(x, new_model.data.train_x)
    _, ref_losses = ref_model.outputs_losses_train(x, None)
    ref_net.zero_grad()
    ref_losses.sum().backward()
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(new_model)
            assert new_model.fused_core._shape == (new_net.layers[0][0].in_features, 100, 5, 2)
            _, new_losses = new_model.outputs_losses_train(x, None)
            new_net.zero_grad()
            new_losses.sum().backward()
            np.testing.assert_allclose(new_losses.detach().numpy(), ref_losses.detach().numpy(), rtol=5e-5)
            for (n, p), (_, q) in zip(ref_net.named_parameters(), new_net.named_parameters()):
                rg = torch.zeros_like(p) if p.grad is None else p.grad
                ng = torch.zeros_like(q) if q.grad is None else q.grad
                assert (ng - rg).norm() <= 5e-5 * max(float(rg.norm()), 1e-9), (n, float((ng - rg).norm()), float(rg.norm()))
            ref_hist, new_hist = _run(ref_model, 3), _run(new_model, 3)
            np.testing.assert_allclose(new_hist, ref_hist, rtol=5e-3)
    finally:
        fused.FusedPDEResidual.__init__ = orig
