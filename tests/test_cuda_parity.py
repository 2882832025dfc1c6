"""Parity of the CUDA path (through the C ABI, libpinn_b200.so) against reference-generated golden
vectors and against the oracle on seeded inputs.  Tolerance: 1e-5 relative (BASELINE.json north_star),
measured as rel-L2 and max-abs/max-abs."""
import numpy as np
import pytest
import torch

from util import GOLDEN_NAMES, Golden, assert_parity, rel_l2

pytestmark = pytest.mark.gpu


def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _to(dev, *ts):
    return [None if t is None else t.to(dev).contiguous() for t in ts]


IMPLS = ["ffma", "tc", "tc2"]


def _select_impl(impl, desc):
    """Both kernels (FP32 FFMA, tcgen05 bf16x3) implement the same C-ABI call; run the parity suite on each."""
    from pinnacle_b200 import capi
    if impl in ("tc", "tc2") and not capi.has_tensor_core_kernel(desc):
        pytest.skip("no tensor-core kernel instantiated for this jet signature")
    capi.set_impl({"tc": capi.IMPL_TC, "tc2": capi.IMPL_TC_PIPELINED, "ffma": capi.IMPL_FFMA}[impl])


@pytest.fixture(autouse=True)
def _reset_impl():
    yield
    from pinnacle_b200 import capi
    capi.set_impl(capi.IMPL_AUTO)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_fwd_bwd_matches_reference_golden(name, impl):
    from pinnacle_b200 import capi
    dev = _cuda()
    G = Golden(name)
    desc = G.desc
    _select_impl(impl, desc)
    flat, x, aux = _to(dev, *G.inputs(torch.float32))
    S, P = desc.n_pde, desc.n_params
    seeds = torch.eye(S, device=dev)
    out = capi.fwd_bwd(desc, flat, x, aux, 1.0 / x.shape[0], seeds)
    assert capi.last_launch_count() == 3
    grads = out[:S * P].view(S, P).cpu().numpy()
    losses = out[S * P:].cpu().numpy()
    loss_f, resid = capi.fwd(desc, flat, x, aux, 1.0 / x.shape[0], want_residual=True)
    assert_parity(resid.cpu().numpy(), G.g["resid64"], "residual vs reference fp64")
    assert_parity(resid.cpu().numpy(), G.g["resid32"], "residual vs reference fp32")
    assert_parity(losses, G.g["loss64"], "loss (fwd_bwd)")
    assert_parity(loss_f.cpu().numpy(), G.g["loss64"], "loss (fwd)")
    assert_parity(grads[:, G.idx], G.g["grad64"], "grad vs reference fp64")
    assert_parity(grads[:, G.idx], G.g["grad32"], "grad vs reference fp32")
    assert_parity(grads.astype(np.float64) @ G.probes.T, G.g["gproj64"], "grad projections")
    assert_parity(np.linalg.norm(grads.astype(np.float64), axis=1), G.g["gnorm64"], "grad norm")


@pytest.mark.parametrize("name,n", [("Poisson2D_Classic", 1), ("Poisson2D_Classic", 32), ("Poisson2D_Classic", 33),
                                    ("Burgers1D", 4099), ("NS2D_LidDriven", 5000), ("KuramotoSivashinskyEquation", 4800),
                                    ("Heat2D_Multiscale", 9001), ("Poisson3D_ComplexGeometry", 2500), ("HeatND", 1111)])
@pytest.mark.parametrize("impl", IMPLS)
def test_matches_oracle_ragged_sizes(name, n, impl):
    """Many tiles per CTA, ragged tails, single point; trained-like weights (x1.5, saturating tanh)."""
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi, problems
    dev = _cuda()
    G = Golden(name)
    desc = G.desc
    _select_impl(impl, desc)
    spec = problems.make(name) if desc.n_aux == 0 or name in ("Poisson3D_ComplexGeometry", "HeatND") else None
    rng = np.random.RandomState(n)
    flat64 = G.flat64 * 1.2
    if spec is not None:
        x64 = spec.sample(n, seed=n, dtype=np.float32).astype(np.float64)
        aux64 = spec.aux(torch.as_tensor(x64))
    else:
        x64 = G.g["x"][rng.randint(0, G.g["x"].shape[0], n)]
        aux64 = None
    seeds = torch.as_tensor(rng.standard_normal((2, desc.n_pde)))
    ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat64), torch.as_tensor(x64), aux64, seeds=seeds)
    flat, x, aux, sd = _to(dev, torch.as_tensor(flat64).float(), torch.as_tensor(x64).float(),
                           None if aux64 is None else aux64.float(), seeds.float())
    out = capi.fwd_bwd(desc, flat, x, aux, 1.0 / n, sd)
    P = desc.n_params
    assert_parity(out[2 * P:].cpu().numpy(), ol.numpy(), "loss")
    assert_parity(out[:2 * P].view(2, P).cpu().numpy(), og.numpy(), "grad")
    _, resid = capi.fwd(desc, flat, x, aux, 1.0 / n, want_residual=True)
    assert_parity(resid.cpu().numpy(), orr.numpy(), "residual")


@pytest.mark.parametrize("impl", IMPLS)
def test_bitwise_deterministic(impl):
    from pinnacle_b200 import capi, problems
    dev = _cuda()
    spec = problems.make("Poisson2D_Classic")
    _select_impl(impl, spec.desc)
    G = Golden("Poisson2D_Classic")
    x = torch.as_tensor(spec.sample(70001, seed=3)).to(dev)
    flat = torch.as_tensor(G.flat64).float().to(dev)
    seeds = torch.ones(1, 1, device=dev)
    a = capi.fwd_bwd(spec.desc, flat, x, None, 1.0 / x.shape[0], seeds).clone()
    b = capi.fwd_bwd(spec.desc, flat, x, None, 1.0 / x.shape[0], seeds).clone()
    assert torch.equal(a, b)


@pytest.mark.parametrize("impl", IMPLS)
def test_repeated_runs_stay_bitwise_identical(impl):
    """Poor man's race check (compute-sanitizer is closed on this pool): 12 repeats x 3 problem shapes x ragged sizes."""
    from pinnacle_b200 import capi, problems
    dev = _cuda()
    for name, n in (("NS2D_LidDriven", 9973), ("Heat2D_Multiscale", 20011), ("KuramotoSivashinskyEquation", 4801)):
        spec = problems.make(name)
        _select_impl(impl, spec.desc)
        x = torch.as_tensor(spec.sample(n, seed=n)).to(dev)
        flat = torch.as_tensor(Golden(name).flat64).float().to(dev)
        seeds = torch.ones(2, spec.desc.n_pde, device=dev)
        seeds[1] *= 0.5
        first = capi.fwd_bwd(spec.desc, flat, x, None, 1.0 / n, seeds).clone()
        assert torch.isfinite(first).all()
        for _ in range(12):
            assert torch.equal(capi.fwd_bwd(spec.desc, flat, x, None, 1.0 / n, seeds), first)


@pytest.mark.parametrize("name", ["Poisson2D_Classic", "Heat2D_Multiscale", "NS2D_LidDriven"])
def test_full_size_properties(name):
    """At BASELINE.json's 1M-point size the oracle is out of reach; check size-independent properties:
    (1) shard additivity: out(all rows) == out(first half) + out(second half) at the same 1/N;
    (2) linearity in the seed;  (3) permutation invariance of the point order;  (4) sub-sample parity
    of the residual rows against the oracle."""
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi, problems
    dev = _cuda()
    spec = problems.make(name)
    desc = spec.desc
    G = Golden(name)
    n = 1 << 20
    xh = spec.sample(n, seed=11)
    x = torch.as_tensor(xh).to(dev)
    flat = torch.as_tensor(G.flat64).float().to(dev)
    P, K = desc.n_params, desc.n_pde
    s1 = torch.ones(1, K, device=dev)
    full = capi.fwd_bwd(desc, flat, x, None, 1.0 / n, s1).clone()
    h = n // 2 + 7
    a = capi.fwd_bwd(desc, flat, x[:h].contiguous(), None, 1.0 / n, s1).clone()
    b = capi.fwd_bwd(desc, flat, x[h:].contiguous(), None, 1.0 / n, s1).clone()
    assert_parity((a + b).cpu().numpy(), full.cpu().numpy(), "shard additivity")
    s2 = torch.as_tensor(np.random.RandomState(0).standard_normal((1, K))).float().to(dev)
    g2 = capi.fwd_bwd(desc, flat, x, None, 1.0 / n, s2)[:P].clone()
    g12 = capi.fwd_bwd(desc, flat, x, None, 1.0 / n, s1 + s2)[:P].clone()
    assert_parity(g12.cpu().numpy(), (full[:P] + g2).cpu().numpy(), "seed linearity")
    perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    pp = capi.fwd_bwd(desc, flat, x[perm].contiguous(), None, 1.0 / n, s1)
    assert_parity(pp.cpu().numpy(), full.cpu().numpy(), "permutation invariance")
    _, resid = capi.fwd(desc, flat, x, None, 1.0 / n, want_residual=True)
    sub = np.random.RandomState(5).choice(n, 512, replace=False)
    _, _, orr = O.losses_and_grads(desc, torch.as_tensor(G.flat64), torch.as_tensor(xh[sub].astype(np.float64)), None)
    assert_parity(resid[torch.as_tensor(sub, device=dev)].cpu().numpy(), orr.numpy(), "residual sub-sample vs oracle fp64")
    assert abs(float((resid.double() ** 2).sum(0).cpu()[0]) / n - float(full[P].cpu())) <= 1e-5 * float(full[P].cpu())


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("L", [2, 3, 7])
def test_other_depths_match_oracle(L, impl):
    """--hidden-layers 100*L for L != 5 (benchmark.py:65): depth is a run-time parameter of both kernels."""
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi, problems
    from golden import common
    dev = _cuda()
    spec = problems.make("Burgers1D")
    desc = spec.desc.with_net(100, L)
    _select_impl(impl, desc)
    flat64 = common.golden_params(desc.layer_shapes(), 11 + L)
    x64 = spec.sample(1500, seed=L, dtype=np.float32).astype(np.float64)
    ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat64), torch.as_tensor(x64), None)
    out = capi.fwd_bwd(desc, torch.as_tensor(flat64).float().to(dev), torch.as_tensor(x64).float().to(dev), None, 1.0 / 1500,
                       torch.ones(1, 1, device=dev))
    assert_parity(out[desc.n_params:].cpu().numpy(), ol.numpy(), "loss")
    assert_parity(out[:desc.n_params].cpu().numpy(), og[0].numpy(), "grad")


@pytest.mark.parametrize("d,m,comp,n", [(2, 1, 0, 129), (2, 3, 2, 1), (3, 2, 1, 2049), (3, 3, 0, 513), (1, 1, 0, 77), (5, 1, 0, 300), (6, 1, 0, 300)])
def test_value_bc_kind_matches_oracle(d, m, comp, n):
    """Fused value-type boundary / initial / point-set term (kind VALUE) vs the oracle restatement of DirichletBC.error."""
    from oracle import autograd_ref as O
    from golden import common
    from pinnacle_b200 import capi, kinds
    dev = _cuda()
    desc = kinds.value_desc(d, m, comp)
    rs = np.random.RandomState(n)
    flat = common.golden_params(desc.layer_shapes(), 5)
    x = rs.uniform(-1, 1, (n, d)).astype(np.float32).astype(np.float64)
    # targets offset so a single-row term (PointSetBC at one point, ns.py:196-201) is not a u ~ g cancellation
    tgt = (rs.standard_normal((n, 1)) + 2.0).astype(np.float32).astype(np.float64)
    ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat), torch.as_tensor(x), torch.as_tensor(tgt))
    out = capi.fwd_bwd(desc, torch.as_tensor(flat).float().to(dev), torch.as_tensor(x).float().to(dev), torch.as_tensor(tgt).float().to(dev),
                       1.0 / n, torch.ones(1, 1, device=dev))
    assert_parity(out[desc.n_params:].cpu().numpy(), ol.numpy(), "loss")
    assert_parity(out[:desc.n_params].cpu().numpy(), og[0].numpy(), "grad")


@pytest.mark.parametrize("d,with_beta,n", [(2, False, 129), (2, True, 1), (2, True, 2049), (3, False, 513), (3, True, 8193), (6, False, 2049), (6, True, 33)])
def test_flux_bc_kind_matches_oracle(d, with_beta, n):
    """Fused flux-type boundary term (kind FLUX: NeumannBC.error / affine RobinBC.error) vs the oracle."""
    from oracle import autograd_ref as O
    from test_host_math import _flux_case
    from pinnacle_b200 import capi
    dev = _cuda()
    desc, flat, x, aux = _flux_case(d, with_beta, 100 * d + n, n=n)
    ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat).double(), torch.as_tensor(x).double(), torch.as_tensor(aux).double())
    out = capi.fwd_bwd(desc, torch.as_tensor(flat).to(dev), torch.as_tensor(x).to(dev), torch.as_tensor(aux).to(dev), 1.0 / n,
                       torch.ones(1, 1, device=dev))
    assert_parity(out[desc.n_params:].cpu().numpy(), ol.numpy(), "loss")
    assert_parity(out[:desc.n_params].cpu().numpy(), og[0].numpy(), "grad")
    loss, resid = capi.fwd(desc, torch.as_tensor(flat).to(dev), torch.as_tensor(x).to(dev), torch.as_tensor(aux).to(dev), 1.0 / n, want_residual=True)
    assert_parity(resid.cpu().numpy(), orr.numpy(), "residual")


def test_errors_are_loud():
    from pinnacle_b200 import capi, kinds
    dev = _cuda()
    bad = kinds.KindDesc(kinds.KIND_KS, 2, 1, (2, 2))          # KS needs a 4th-order jet
    with pytest.raises(capi.FusedOpError):
        capi.check_supported(bad)
    G = Golden("Poisson2D_Classic")
    flat, x, _ = G.inputs(torch.float32)
    with pytest.raises(capi.FusedOpError):                       # CPU tensors: no fallback
        capi.fwd(G.desc, flat, x, None, 1.0)
    with pytest.raises(capi.FusedOpError):                       # a hidden width no kernel is instantiated for (100, 64, 128 are)
        capi.check_supported(G.desc.with_net(80, 5))
    with pytest.raises(capi.FusedOpError):
        capi.fwd(G.desc, flat.to(dev)[:-1].contiguous(), x.to(dev), None, 1.0)


@pytest.mark.parametrize("name", ["Poisson2D_Classic", "NS2D_LidDriven", "KuramotoSivashinskyEquation", "Heat2D_Multiscale", "PoissonND"])
@pytest.mark.parametrize("impl", ["ffma", "tc", "tc2"])
def test_linear_seed_rows_match_the_gradient_of_the_residual_sums(name, impl):
    """pinn_fwd_bwd_linear: row s = d(sum_k seeds[s,k] sum_rows r_k)/dtheta (the NTK statistics of src/optimizer/ntk.py:24-58)
    against the reference algorithm's autograd gradient of the same sums, on the golden points and weights (fp64 oracle)."""
    from oracle import autograd_ref as O
    from pinnacle_b200 import capi
    dev = _cuda()
    G = Golden(name)
    flat, x, aux = G.inputs(torch.float64)
    p = flat.clone().requires_grad_(True)
    _, res, _ = O.pde_losses(G.desc, p, x, aux)
    rs = np.random.RandomState(3)
    seeds = np.vstack([np.ones(G.desc.n_pde), rs.standard_normal(G.desc.n_pde)])
    refs = [torch.autograd.grad((res * torch.as_tensor(sd)).sum(), p, retain_graph=True)[0].numpy() for sd in seeds]
    capi.set_impl({"ffma": capi.IMPL_FFMA, "tc": capi.IMPL_TC, "tc2": capi.IMPL_TC_PIPELINED}[impl])
    try:
        if impl != "ffma" and not capi.has_tensor_core_kernel(G.desc):
            pytest.skip("no tensor-core kernel for this configuration")
        out = capi.fwd_bwd_linear(G.desc, flat.float().to(dev), x.float().to(dev), None if aux is None else aux.float().to(dev),
                                  torch.as_tensor(seeds, dtype=torch.float32).to(dev))
    finally:
        capi.set_impl(capi.IMPL_AUTO)
    P = G.desc.n_params
    for s, ref in enumerate(refs):
        assert_parity(out[s * P:(s + 1) * P].cpu().numpy(), ref, f"linear row {s}")
