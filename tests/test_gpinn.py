"""gPINN (src/pde/baseclass.py:151-180 ``use_gepinn``, benchmark.py --method gepinn) without a GPU: the oracle
and the DEVICE math (jet_math.cuh compiled for the host) against reference-generated vectors, and the drop-in on a real
``dde.Model`` with the oracle standing in for the kernels."""
import ctypes

import numpy as np
import pytest
import torch

import refshim
from util import GpinnGolden, assert_parity, oracle_backend

CASES = ["Poisson2D_Classic", "Burgers1D"]
GOLDEN_CASES = CASES + ["Wave1D", "Helmholtz2D", "PoissonBoltzmann2D", "Poisson2D_ManyArea", "Poisson1D"]      # tests/golden/gepinn_<Class>.npz
needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not available")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_gepinn(name):
    from oracle import autograd_ref as O
    G = GpinnGolden(name)
    # (Poisson2D_ManyArea: the fused path keeps the coefficient columns in float32 as the reference's float32 run does,
    # poisson.py:251; the reference's float64 run holds them in double: 2e-8)
    tol64 = 2e-7 if name == "Poisson2D_ManyArea" else 1e-9
    for dtype, tag, tol in ((torch.float64, "64", tol64), (torch.float32, "32", 2e-5)):
        flat, x = G.embedded(dtype)
        losses, grads, res = O.losses_and_grads(G.desc, flat, x, G.aux(x), seeds=np.eye(G.desc.n_pde))
        grads = G.back(grads.numpy())
        assert_parity(res.numpy(), G.g[f"resid{tag}"], "residual", tol)
        assert_parity(losses.numpy(), G.g[f"loss{tag}"], "loss", tol)
        assert_parity(grads[:, G.idx], G.g[f"grad{tag}"], "grad", tol)
        assert_parity(grads @ G.probes.T, G.g[f"gproj{tag}"], "grad projections", tol)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_device_math_matches_reference_gepinn(harness_lib, name):
    """Order-3 jets along e0, e1, e0+e1, e0-e1, the polarisation formulas for u_xy, u_xxy, u_xyy and their hand adjoints."""
    G = GpinnGolden(name)
    desc = G.desc
    flat, x = G.embedded(torch.float32)
    aux = G.aux(x)
    auxn = None if aux is None else np.ascontiguousarray(aux.numpy())
    flat, x = flat.numpy(), x.numpy()
    n, P, S = x.shape[0], desc.n_params, desc.n_pde
    seeds = np.eye(S, dtype=np.float32)
    grad = np.zeros((S, P))
    loss = np.zeros(3)
    resid = np.zeros((n, S), np.float32)
    cd = desc.to_c()
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = harness_lib.harness_run(ctypes.byref(cd), vp(flat), vp(x), vp(auxn), ctypes.c_longlong(n), ctypes.c_double(1.0 / n), vp(seeds), S,
                                 vp(grad), vp(loss), vp(resid))
    assert rc == 0
    grad = G.back(grad)
    assert_parity(resid, G.g["resid64"], "residual")
    assert_parity(loss[:S], G.g["loss64"], "loss")
    assert_parity(grad[:, G.idx], G.g["grad64"], "grad")
    assert_parity(grad @ G.probes.T, G.g["gproj64"], "grad projections")


@needs_ref
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_patched_gepinn_model_matches_unpatched_reference_step(name):
    """benchmark.py:84-122 with --method gepinn: pde.use_gepinn(), create_model, compile, enable_fused -- same losses (three PDE
    terms + the boundary terms), same parameter gradients, one optimizer step through the patched train_step."""
    dde, _ = refshim.load()
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    dde.config.set_random_seed(0)
    pde = refshim.construct(name)
    pde.use_gepinn()
    pde.training_points(domain=96, boundary=32, **({"initial": 32} if hasattr(pde, "num_initial_points") else {}))
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [1], "tanh", "Glorot normal").float()
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    model = pde.create_model(net)
    w = np.linspace(1.0, 2.0, pde.num_loss)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=w)
    net.auxiliary_vars = None
    x = model.data.train_x
    _, ref_losses = model.outputs_losses_train(x, None)
    n_terms = 1 + pde.input_dim
    assert ref_losses.shape[0] == n_terms + len(model.data.bcs)
    net.zero_grad()
    ref_losses.sum().backward()
    ref_grads = [p.grad.clone() for p in net.parameters()]
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            assert model.fused_core.op.desc.n_pde == n_terms and model.fused_self_check < 1e-5
            _, new_losses = model.outputs_losses_train(x, None)
            net.zero_grad()
            new_losses.sum().backward()
            np.testing.assert_allclose(new_losses.detach().numpy(), ref_losses.detach().numpy(), rtol=2e-5)
            for p, rg in zip(net.parameters(), ref_grads):
                assert (p.grad - rg).norm() <= 2e-5 * max(float(rg.norm()), 1e-12), (name, float((p.grad - rg).norm()), float(rg.norm()))
            model._train_step(x, None, None)
            assert model.opt.losses.shape[0] == pde.num_loss and torch.isfinite(model.opt.losses[:n_terms]).all()   # (a boundary term without rows is nan in the reference too)
    finally:
        fused.FusedPDEResidual.__init__ = orig


@needs_ref
def test_gepinn_outside_the_covered_problems_raises():
    from pinnacle_b200 import problems
    for name in ("Heat2D_Multiscale", "NS2D_LidDriven"):
        pde = refshim.construct(name)
        pde.use_gepinn()
        with pytest.raises(problems.UnsupportedProblem, match="gPINN"):
            problems.describe_reference(pde)
