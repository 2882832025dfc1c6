"""Host-side logic of the drop-in boundary on CPU (oracle injected as the compute backend):
autograd semantics of the fused loss node, the reference-free solver, and -- where the reference tree is
mounted -- the patched dde.Model against the unpatched reference step."""
import numpy as np
import pytest
import torch

import refshim
from util import oracle_backend

needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


def _net(desc, seed=0):
    from pinnacle_b200.nn import FNN
    torch.manual_seed(seed)
    net = FNN([desc.d] + [100] * 5 + [desc.m])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    return net


def test_loss_node_supports_repeated_seeded_backward():
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    from pinnacle_b200.fused import FusedPDEResidual
    spec = problems.make("NS2D_LidDriven")
    net = _net(spec.desc)
    params = [t for lin in net.linears for t in (lin.weight, lin.bias)]
    x = spec.sample(50, seed=2)
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    ol, og, _ = O.losses_and_grads(spec.desc, flat, torch.as_tensor(x), None)
    for policy in ("lazy", "eager"):
        with oracle_backend():
            op = FusedPDEResidual(spec, policy=policy, device="cpu", check_library=False).set_points(x)
            losses = op.losses(params)
            np.testing.assert_allclose(losses.detach().numpy(), ol.numpy(), rtol=1e-6)
            for k in range(3):                       # per-term backward accumulates like autograd (lr_adaptor.py:36,49)
                net.zero_grad()
                losses[k].backward(retain_graph=True)
                g = torch.cat([p.grad.reshape(-1) for p in params])
                assert (g - og[k]).norm() <= 1e-5 * og[k].norm()
            net.zero_grad()
            losses.sum().backward(retain_graph=True)
            losses.sum().backward()                  # second call accumulates
            g = torch.cat([p.grad.reshape(-1) for p in params])
            assert (g - 2 * og.sum(0)).norm() <= 1e-5 * (2 * og.sum(0)).norm()


def test_forward_only_when_params_frozen():
    from pinnacle_b200 import problems
    from pinnacle_b200.fused import FusedPDEResidual
    spec = problems.make("Poisson2D_Classic")
    net = _net(spec.desc)
    net.requires_grad_(False)                        # what Model._outputs_losses does for _test() (model.py:490)
    params = [t for lin in net.linears for t in (lin.weight, lin.bias)]
    with oracle_backend():
        op = FusedPDEResidual(spec, device="cpu", check_library=False).set_points(spec.sample(20))
        losses = op.losses(params)
    assert not losses.requires_grad
    assert op.stats["fwd_calls"] == 1 and op.stats["fwd_bwd_calls"] == 0


def test_solver_adam_steps_follow_oracle_training():
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    from pinnacle_b200.solver import FusedPDESolver
    spec = problems.make("Burgers1D")
    x = spec.sample(64, seed=3)
    x_bc = spec.sample(16, seed=4)
    x_bc[:, 1] = 0.0
    target = torch.as_tensor(np.sin(-np.pi * x_bc[:, 0:1]))
    net_a, net_b = _net(spec.desc, 1), _net(spec.desc, 1)
    w = np.array([1.0, 3.0])
    with oracle_backend():
        solver = FusedPDESolver(spec, net_a, x, x_bc, [lambda xb, yb: yb[:, 0:1] - target], device="cpu")
        solver.core.op  # constructed without touching the CUDA library?  (check_library happens inside)
        solver.compile(torch.optim.Adam(net_a.parameters(), 1e-3), loss_weights=w)
        hist_a = []
        for _ in range(3):
            solver.train_step()
            hist_a.append(solver.opt.losses.detach().numpy().copy())
    opt = torch.optim.Adam(net_b.parameters(), 1e-3)
    hist_b = []
    for _ in range(3):
        opt.zero_grad()
        flat = torch.cat([p.reshape(-1) for p in net_b.parameters()])
        losses, _, _ = O.pde_losses(spec.desc, flat, torch.as_tensor(x), None)
        bc = torch.mean(torch.square(net_b(torch.as_tensor(x_bc))[:, 0:1] - target))
        all_l = torch.cat([losses, bc.reshape(1)]) * torch.as_tensor(w, dtype=torch.float32)
        all_l.sum().backward()
        opt.step()
        hist_b.append(all_l.detach().numpy().copy())
    np.testing.assert_allclose(np.array(hist_a), np.array(hist_b), rtol=2e-4)


@needs_ref
@pytest.mark.parametrize("name", ["Burgers1D", "Poisson2D_Classic", "NS2D_LidDriven", "Wave1D", "Heat2D_ComplexGeometry", "Poisson2D_ManyArea",
                                  "Wave2D_LongTime", "Burgers2D", "HeatND", "PoissonInv", "HeatInv"])
def test_patched_dde_model_matches_unpatched_reference_step(name):
    """Same weights, same points: losses and p.grad of the reference closure vs the patched closure."""
    import copy
    dde, _ = refshim.load()
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    dde.config.set_random_seed(0)
    pde = refshim.construct(name)
    pde.training_points(domain=96, boundary=32, **({"initial": 32} if hasattr(pde, "num_initial_points") else {}))
    pde.num_test_points = 64
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    model = pde.create_model(net)
    w = np.linspace(1.0, 2.0, pde.num_loss)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=w)
    net.auxiliary_vars = None
    x = model.data.train_x
    # reference closure
    _, ref_losses = model.outputs_losses_train(x, None)
    net.zero_grad()
    ref_losses.sum().backward()
    ref_grads = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    ref_test = model._outputs_losses(False, model.data.test_x, None, None)
    with oracle_backend():
        from pinnacle_b200 import capi
        from pinnacle_b200 import fused
        orig = fused.FusedPDEResidual.__init__

        def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
            orig(self, spec, H, L, policy, "cpu", group, False)

        fused.FusedPDEResidual.__init__ = init_cpu
        try:
            enable_fused(model)
            # patch-time self-check: the live pde() closure vs the fused residual on 64 training rows
            assert model.fused_self_check is not None and model.fused_self_check < 1e-5, model.fused_self_check
            _, new_losses = model.outputs_losses_train(x, None)
            net.zero_grad()
            new_losses.sum().backward()
            np.testing.assert_allclose(new_losses.detach().numpy(), ref_losses.detach().numpy(), rtol=2e-5)
            for p, rg in zip(net.parameters(), ref_grads):
                ng = torch.zeros_like(p) if p.grad is None else p.grad
                rg = torch.zeros_like(p) if rg is None else rg
                assert (ng - rg).norm() <= 2e-5 * max(rg.norm(), 1e-12), (name, (ng - rg).norm(), rg.norm())
            # _test() path: frozen params, losses on the test rows, outputs for every row
            new_test = model._outputs_losses(False, model.data.test_x, None, None)
            np.testing.assert_allclose(new_test[1], ref_test[1], rtol=2e-5)
            np.testing.assert_allclose(new_test[0], ref_test[0], rtol=1e-5, atol=1e-6)
            # one optimizer step through the patched train_step keeps the reference's protocol
            model._train_step(x, None, None)
            assert model.opt.losses.shape[0] == pde.num_loss
        finally:
            fused.FusedPDEResidual.__init__ = orig


@needs_ref
def test_edited_pde_under_a_known_class_name_is_caught_at_patch_time():
    """describe_reference maps a problem by its class name; a model whose live pde() closure computes something else (a
    subclass, an edited residual) must not silently train on the canonical formula: enable_fused compares the two on a few
    training rows and raises."""
    dde, _ = refshim.load()
    from pinnacle_b200 import capi, fused
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    pde = refshim.construct("Burgers1D")
    pde.training_points(domain=96, boundary=32, initial=32)
    net = dde.nn.FNN([2] + [100] * 5 + [1], "tanh", "Glorot normal").float()
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    model = pde.create_model(net)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
    canonical = model.data.pde
    model.data.pde = lambda x, u: canonical(x, u) + 1e-2 * u           # "the same" Burgers1D with an extra reaction term
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            with pytest.raises(capi.FusedOpError, match="self-check"):
                enable_fused(model)
            enable_fused(model, self_check=0)                           # explicit opt-out
            assert model.fused_self_check is None
    finally:
        fused.FusedPDEResidual.__init__ = orig


@needs_ref
def test_unsupported_problems_raise():
    from pinnacle_b200 import problems
    dde, _ = refshim.load()
    pde = refshim.construct("Heat2D_Multiscale")      # gPINN in three inputs: not covered (tests/test_gpinn.py has the covered cases)
    pde.use_gepinn()
    with pytest.raises(problems.UnsupportedProblem):
        problems.describe_reference(pde)

    class Strange:
        pde = staticmethod(lambda x, u: u)
        bbox = [0, 1]
    with pytest.raises(problems.UnsupportedProblem):
        problems.describe_reference(Strange())


@needs_ref
@pytest.mark.parametrize("name,n_flux", [("Wave1D", 1), ("Heat2D_ComplexGeometry", 3), ("Poisson2D_ManyArea", 1), ("HeatND", 1)])
def test_flux_boundary_terms_take_the_fused_path(name, n_flux):
    """NeumannBC and affine RobinBC terms (m = 1) are evaluated by the fused op (kind FLUX), not by stock autograd --
    including the 6-dimensional Neumann term of HeatND."""
    dde, _ = refshim.load()
    from pinnacle_b200 import fused, kinds
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    pde = refshim.construct(name)
    pde.training_points(domain=96, boundary=64, **({"initial": 32} if hasattr(pde, "num_initial_points") else {}))
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
    model = pde.create_model(net)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
    net.auxiliary_vars = None
    _, ref_losses = model.outputs_losses_train(model.data.train_x, None)
    made = []
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        made.append(spec.desc.kind)
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            _, losses = model.outputs_losses_train(model.data.train_x, None)
    finally:
        fused.FusedPDEResidual.__init__ = orig
    assert made.count(kinds.KIND_FLUX) == n_flux
    np.testing.assert_allclose(losses.detach().numpy(), ref_losses.detach().numpy(), rtol=2e-5)


def test_term_gradient_rows_equal_per_term_autograd_backward():
    """FusedPDESolver.term_gradient_rows (host logic of the autograd-free optimizers, oracle as compute backend): row t is
    what losses[t].backward() leaves in p.grad, the loss vector is the unweighted one."""
    from pinnacle_b200 import fused, problems
    from pinnacle_b200.solver import FusedPDESolver, ValueBC
    spec = problems.make("NS2D_LidDriven")
    x = spec.sample(160, seed=1)
    xb = spec.sample(48, seed=2)
    terms = [ValueBC(0, 32, 0, np.ones((32, 1), np.float32)), ValueBC(32, 48, 2, np.zeros((16, 1), np.float32))]
    net = _net(spec.desc)
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
            _, losses = s.outputs_losses_train()
            np.testing.assert_allclose(L.numpy(), losses.detach().numpy(), rtol=1e-5)
            assert G.shape == (5, spec.desc.n_params)
            for t in range(5):
                net.zero_grad()
                losses[t].backward(retain_graph=True)
                g = torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1) for p in net.parameters()])
                assert (G[t] - g).norm() <= 2e-5 * g.norm()
            G1, _ = s.term_gradient_rows(flat, torch.tensor([[1.0, 2.0, 0.5]]))
            assert (G1[0] - (G[0] + 2 * G[1] + 0.5 * G[2])).norm() <= 2e-5 * G1[0].norm()
    finally:
        fused.FusedPDEResidual.__init__ = orig


@needs_ref
def test_every_boundary_term_of_burgers2d_and_heatnd_is_fused():
    """Burgers2D: 2 IC + 4 PeriodicBC terms; HeatND: IC + 6-dimensional NeumannBC -- no stock-autograd BC evaluation is left
    (the net is never called on the BC rows in training mode)."""
    dde, _ = refshim.load()
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        for name in ("Burgers2D", "HeatND"):
            torch.manual_seed(0)
            pde = refshim.construct(name)
            pde.training_points(domain=64, boundary=48, initial=32)
            net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
            model = pde.create_model(net)
            model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
            net.auxiliary_vars = None
            calls = []
            fwd = net.forward
            net.forward = lambda x, fwd=fwd: (calls.append(x.shape), fwd(x))[1]
            with oracle_backend():
                enable_fused(model)
                assert calls == [torch.Size([64, pde.input_dim])]      # the patch-time self-check's reference evaluation, once
                calls.clear()
                _, losses = model.outputs_losses_train(model.data.train_x, None)
                losses.sum().backward()
            assert calls == [], (name, calls)
            assert losses.shape[0] == pde.num_loss and torch.isfinite(losses).all()
    finally:
        fused.FusedPDEResidual.__init__ = orig


def test_resampled_points_of_the_same_shape_are_restaged():
    """The point cache is keyed on the host array's address; the staged array is kept alive so that a freshly resampled
    array of the same shape (PDEPointResampler) can never alias a freed one and be mistaken for the staged set."""
    import gc
    from pinnacle_b200 import fused, problems
    from pinnacle_b200.solver import FusedPDESolver
    spec = problems.make("Poisson2D_Classic")
    net = _net(spec.desc)
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            s = FusedPDESolver(spec, net, spec.sample(64, seed=1), device="cpu").compile(torch.optim.Adam(net.parameters(), 1e-3))
            seen = set()
            for k in range(6):
                x = spec.sample(64, seed=10 + k)           # same shape every time, previous array unreferenced by the caller
                _, losses = s.outputs_losses_train(x)
                ref = float(losses[0])
                gen = s.core.generation
                _, again = s.outputs_losses_train(x)       # same array: cache hit
                assert s.core.generation == gen and float(again[0]) == ref
                assert s.core._key_ref is x
                seen.add(round(ref, 12))
                del x
                gc.collect()
            assert len(seen) == 6                           # six different point sets gave six different losses
    finally:
        fused.FusedPDEResidual.__init__ = orig
