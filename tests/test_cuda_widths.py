"""Hidden widths other than 100 (benchmark.py:65 ``--hidden-layers``, src/utils/args.py:4-12): H = 64 and H = 128 on the FP32 FFMA
kernel -- residual terms and fused boundary terms against the oracle (reference algorithm, fp64), through the C ABI and through
the solver front end.  Widths that are not instantiated raise at patch time, never fall back."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(desc, seed, scale=1.25):
    from golden import common
    return common.golden_params(desc.layer_shapes(), seed)


_TABLE = np.array([[0.0, 0.0, 1.0], [1.0, 1.0, 2.0], [0.0, 1.0, 1.5], [1.0, 0.0, 1.2]])
_MAKE_ARGS = {"Wave2D_Heterogeneous": (_TABLE,), "Heat2D_VaryingCoef": (_TABLE,)}


def _spec(name):
    from pinnacle_b200 import problems
    return problems.BUILDERS[name](*_MAKE_ARGS.get(name, ()))


def _check_kind(name, H, L, n=1111):
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi
    dev = torch.device("cuda:0")
    spec = _spec(name)
    desc = spec.desc.with_net(H, L)
    assert capi.supported(desc) and capi.selected_impl(desc) == capi.IMPL_FFMA
    flat64 = _params(desc, 11 + H + L)
    x = spec.sample(n, seed=H)
    xt = torch.as_tensor(x)
    aux = spec.aux(xt)
    S = desc.n_pde
    out = capi.fwd_bwd(desc, torch.as_tensor(flat64.astype(np.float32)).to(dev), xt.to(dev), None if aux is None else aux.to(dev),
                       1.0 / x.shape[0], torch.eye(S, device=dev))
    P = desc.n_params
    ol, og, ores = O.losses_and_grads(desc, torch.as_tensor(flat64), xt.double(), None if aux is None else aux.double(), seeds=np.eye(S))
    np.testing.assert_allclose(out[S * P:].cpu().numpy(), ol.numpy(), rtol=1e-5)
    g = out[:S * P].view(S, P).cpu().double().numpy()
    for k in range(S):
        assert np.linalg.norm(g[k] - og[k].numpy()) <= 1e-5 * np.linalg.norm(og[k].numpy()), (name, H, L, k)


@pytest.mark.parametrize("H", [64, 128])
@pytest.mark.parametrize("name", ["Burgers2D", "GrayScottEquation", "Poisson1D", "PoissonND", "HeatND", "NS2D_LongTime", "Wave2D_Heterogeneous",
                                  "PoissonInv", "HeatInv", "Poisson3D_ComplexGeometry", "Heat2D_LongTime", "Heat2D_VaryingCoef"])
def test_every_class_has_kernels_at_widths_64_and_128(name, H):
    """The classes outside the five families of the test below: same bar (1e-5 against the fp64 reference algorithm)."""
    _check_kind(name, H, 5, n=777)


@pytest.mark.parametrize("H", [64, 128])
@pytest.mark.parametrize("d,m,comp", [(3, 2, 1), (3, 3, 2), (1, 1, 0), (5, 1, 0), (6, 1, 0)])
def test_value_terms_of_the_remaining_shapes_at_other_widths(d, m, comp, H):
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi
    from pinnacle_b200.kinds import value_desc
    dev = torch.device("cuda:0")
    desc = value_desc(d, m, comp).with_net(H, 5)
    assert capi.supported(desc)
    rs = np.random.RandomState(d * 10 + m)
    x = rs.uniform(-1, 1, (333, d)).astype(np.float32)
    aux = rs.standard_normal((333, 1)).astype(np.float32)
    flat64 = _params(desc, 5 + H + d)
    out = capi.fwd_bwd(desc, torch.as_tensor(flat64.astype(np.float32)).to(dev), torch.as_tensor(x).to(dev), torch.as_tensor(aux).to(dev),
                       1.0 / 333, torch.ones(1, 1, device=dev))
    P = desc.n_params
    ol, og, _ = O.losses_and_grads(desc, torch.as_tensor(flat64), torch.as_tensor(x).double(), torch.as_tensor(aux).double(), seeds=np.ones((1, 1)))
    np.testing.assert_allclose(out[P:].cpu().numpy(), ol.numpy(), rtol=1e-5)
    assert np.linalg.norm(out[:P].cpu().double().numpy() - og[0].numpy()) <= 1e-5 * np.linalg.norm(og[0].numpy())


@pytest.mark.parametrize("H", [64, 128])
def test_six_dimensional_neumann_term_at_other_widths(H):
    """HeatND's Neumann term (heat.py:296-301): n . grad u - g on the 6-dimensional boundary rows."""
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi
    from pinnacle_b200.kinds import flux_desc
    dev = torch.device("cuda:0")
    desc = flux_desc(6, 1, 0, False).with_net(H, 5)
    assert capi.supported(desc)
    rs = np.random.RandomState(6)
    x = rs.uniform(-1, 1, (300, 6)).astype(np.float32)
    aux = rs.standard_normal((300, desc.n_aux)).astype(np.float32)
    flat64 = _params(desc, 9 + H)
    out = capi.fwd_bwd(desc, torch.as_tensor(flat64.astype(np.float32)).to(dev), torch.as_tensor(x).to(dev), torch.as_tensor(aux).to(dev),
                       1.0 / 300, torch.ones(1, 1, device=dev))
    P = desc.n_params
    ol, og, _ = O.losses_and_grads(desc, torch.as_tensor(flat64), torch.as_tensor(x).double(), torch.as_tensor(aux).double(), seeds=np.ones((1, 1)))
    np.testing.assert_allclose(out[P:].cpu().numpy(), ol.numpy(), rtol=1e-5)
    assert np.linalg.norm(out[:P].cpu().double().numpy() - og[0].numpy()) <= 1e-5 * np.linalg.norm(og[0].numpy())


@pytest.mark.parametrize("H,L", [(64, 5), (128, 5), (64, 3), (128, 8)])
@pytest.mark.parametrize("name", ["Burgers1D", "Poisson2D_Classic", "Heat2D_Multiscale", "NS2D_LidDriven", "KuramotoSivashinskyEquation"])
def test_residual_kinds_at_other_widths_match_oracle(name, H, L):
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi, problems
    dev = torch.device("cuda:0")
    spec = problems.make(name)
    desc = spec.desc.with_net(H, L)
    assert capi.supported(desc) and capi.selected_impl(desc) == capi.IMPL_FFMA
    flat64 = _params(desc, 11 + H + L)
    x = spec.sample(1111, seed=H)
    xt = torch.as_tensor(x)
    aux = spec.aux(xt)
    S = desc.n_pde
    out = capi.fwd_bwd(desc, torch.as_tensor(flat64.astype(np.float32)).to(dev), xt.to(dev), None if aux is None else aux.to(dev),
                       1.0 / x.shape[0], torch.eye(S, device=dev))
    P = desc.n_params
    ol, og, ores = O.losses_and_grads(desc, torch.as_tensor(flat64), xt.double(), None if aux is None else aux.double(), seeds=np.eye(S))
    np.testing.assert_allclose(out[S * P:].cpu().numpy(), ol.numpy(), rtol=1e-5)
    g = out[:S * P].view(S, P).cpu().double().numpy()
    for k in range(S):
        assert np.linalg.norm(g[k] - og[k].numpy()) <= 1e-5 * np.linalg.norm(og[k].numpy()), (name, H, L, k)
    _, res = capi.fwd(desc, torch.as_tensor(flat64.astype(np.float32)).to(dev), xt.to(dev), None if aux is None else aux.to(dev),
                      1.0 / x.shape[0], want_residual=True)
    r = res.cpu().double().numpy()
    assert np.linalg.norm(r - ores.numpy()) <= 1e-5 * np.linalg.norm(ores.numpy())


@pytest.mark.parametrize("H", [64, 128])
def test_solver_with_fused_boundary_terms_at_other_widths(H):
    """Value- and flux-type boundary terms run through the fused kernels of the same width; a few training steps stay finite and
    the per-term losses equal the stock-PyTorch evaluation of the same net."""
    from pinnacle_b200 import problems
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.solver import FluxBC, FusedPDESolver, ValueBC
    dev = torch.device("cuda:0")
    spec = problems.make("Heat2D_Multiscale")
    torch.manual_seed(H)
    net = FNN([3] + [H] * 4 + [1]).to(dev)
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    x = spec.sample(3000, seed=1)
    xb = spec.sample(512, seed=2)
    rs = np.random.RandomState(3)
    vals = rs.standard_normal((256, 1)).astype(np.float32)
    nrm = rs.standard_normal((256, 3)).astype(np.float32)
    g = rs.standard_normal((256, 1)).astype(np.float32)
    terms = [ValueBC(0, 256, 0, vals), FluxBC(256, 512, 0, nrm, g)]
    solver = FusedPDESolver(spec, net, x, xb, terms).compile(torch.optim.Adam(net.parameters(), 1e-3))
    _, losses = solver.outputs_losses_train()
    xbt = torch.as_tensor(xb).to(dev).requires_grad_(True)
    y = net(xbt)
    want_v = torch.mean((y[:256] - torch.as_tensor(vals).to(dev)) ** 2)
    dy = torch.autograd.grad(y.sum(), xbt)[0]
    want_f = torch.mean(((dy[256:] * torch.as_tensor(nrm).to(dev)).sum(1, keepdim=True) - torch.as_tensor(g).to(dev)) ** 2)
    np.testing.assert_allclose(float(losses[1]), float(want_v), rtol=2e-5)
    np.testing.assert_allclose(float(losses[2]), float(want_f), rtol=2e-5)
    for _ in range(5):
        solver.train_step()
    assert torch.isfinite(solver.opt.losses).all()


def test_unsupported_width_raises_at_construction():
    from pinnacle_b200 import capi, problems
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.solver import FusedPDESolver
    spec = problems.make("Burgers1D")
    net = FNN([2] + [80] * 5 + [1]).to("cuda:0")
    with pytest.raises(capi.FusedOpError):
        FusedPDESolver(spec, net, spec.sample(100, seed=0))
