"""The DEVICE math (pinnacle_b200/csrc/jet_math.cuh: tanh Taylor-jet recurrence, its hand-derived adjoint,
the residual functors and their adjoints) compiled for the host and checked against reference-generated
golden vectors -- catches derivation errors without a GPU."""
import ctypes

import numpy as np
import pytest
import torch

from util import GOLDEN_NAMES, Golden, assert_parity


def _run(lib, G, seeds):
    desc = G.desc
    flat, x, aux = G.inputs(torch.float32)
    flat, x = flat.numpy(), np.ascontiguousarray(x.numpy())
    auxn = np.ascontiguousarray(aux.numpy()) if aux is not None else None
    n, S, P = x.shape[0], seeds.shape[0], desc.n_params
    grad = np.zeros((S, P))
    loss = np.zeros(3)
    resid = np.zeros((n, desc.n_pde), np.float32)
    cd = desc.to_c()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = lib.harness_run(ctypes.byref(cd), vp(flat), vp(x), vp(auxn), ctypes.c_longlong(n), ctypes.c_double(1.0 / n),
                         vp(seeds), S, vp(grad), vp(loss), vp(resid))
    assert rc == 0
    return loss[:desc.n_pde], grad, resid


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_device_math_matches_reference(harness_lib, name):
    G = Golden(name)
    seeds = np.eye(G.desc.n_pde, dtype=np.float32)
    loss, grad, resid = _run(harness_lib, G, seeds)
    assert_parity(resid, G.g["resid64"], "residual")
    assert_parity(loss, G.g["loss64"], "loss")
    assert_parity(grad[:, G.idx], G.g["grad64"], "grad")
    assert_parity(grad @ G.probes.T, G.g["gproj64"], "grad projections")
    assert_parity(resid, G.g["resid32"], "residual vs reference fp32")
    assert_parity(grad[:, G.idx], G.g["grad32"], "grad vs reference fp32")


def test_structural_zero_output_bias(harness_lib):
    # residuals made only of derivatives leave the output bias without gradient
    G = Golden("Poisson2D_Classic")
    _, grad, _ = _run(harness_lib, G, np.eye(1, dtype=np.float32))
    assert grad[0, -1] == 0.0


def test_value_kind_matches_oracle(harness_lib):
    """Value-type boundary term (kind VALUE, jet-free signature) of the device math vs the oracle."""
    from oracle import autograd_ref as O
    from golden import common
    from pinnacle_b200 import kinds
    for d, m, comp in ((2, 1, 0), (2, 3, 2), (3, 2, 1)):
        desc = kinds.value_desc(d, m, comp)
        rs = np.random.RandomState(d * 10 + m)
        flat = common.golden_params(desc.layer_shapes(), 5).astype(np.float32)
        x = rs.uniform(-1, 1, (53, d)).astype(np.float32)
        tgt = rs.standard_normal((53, 1)).astype(np.float32)
        ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat).double(), torch.as_tensor(x).double(), torch.as_tensor(tgt).double())
        n, P = 53, desc.n_params
        grad = np.zeros((1, P)); loss = np.zeros(3); resid = np.zeros((n, 1), np.float32)
        cd = desc.to_c()
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        seeds = np.ones((1, 1), np.float32)
        rc = harness_lib.harness_run(ctypes.byref(cd), vp(flat), vp(x), vp(tgt), ctypes.c_longlong(n), ctypes.c_double(1.0 / n), vp(seeds), 1,
                                     vp(grad), vp(loss), vp(resid))
        assert rc == 0
        assert_parity(resid, orr.numpy(), "residual")
        assert_parity(loss[:1], ol.numpy(), "loss")
        assert_parity(grad, og.numpy(), "grad")


def _flux_case(d, with_beta, seed, n=61):
    from golden import common
    from pinnacle_b200 import kinds
    desc = kinds.flux_desc(d, 1, 0, with_beta)
    rs = np.random.RandomState(seed)
    flat = common.golden_params(desc.layer_shapes(), 5).astype(np.float32)
    x = rs.uniform(-1, 1, (n, d)).astype(np.float32)
    nrm = rs.standard_normal((n, d))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    cols = [rs.standard_normal((n, 1)) + 2.0]                  # g, offset so r is not a near-cancellation
    if with_beta:
        cols.append(rs.uniform(0.5, 2.0, (n, 1)))
    aux = np.concatenate(cols + [nrm], axis=1).astype(np.float32)
    return desc, flat, x, np.ascontiguousarray(aux)


def test_flux_kind_matches_oracle(harness_lib):
    """Flux-type boundary term (Neumann / affine Robin: kind FLUX, first-order jets) of the device math vs the oracle."""
    from oracle import autograd_ref as O
    for d, with_beta in ((2, False), (2, True), (3, False), (3, True), (6, False), (6, True)):
        desc, flat, x, aux = _flux_case(d, with_beta, 7 * d + with_beta)
        ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat).double(), torch.as_tensor(x).double(), torch.as_tensor(aux).double())
        n, P = x.shape[0], desc.n_params
        grad = np.zeros((1, P)); loss = np.zeros(3); resid = np.zeros((n, 1), np.float32)
        cd = desc.to_c()
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        seeds = np.ones((1, 1), np.float32)
        rc = harness_lib.harness_run(ctypes.byref(cd), vp(flat), vp(x), vp(aux), ctypes.c_longlong(n), ctypes.c_double(1.0 / n), vp(seeds), 1,
                                     vp(grad), vp(loss), vp(resid))
        assert rc == 0
        assert_parity(resid, orr.numpy(), "residual")
        assert_parity(loss[:1], ol.numpy(), "loss")
        assert_parity(grad, og.numpy(), "grad")


@pytest.mark.parametrize("name", ["Poisson2D_Classic", "Burgers1D", "NS2D_LidDriven", "Heat2D_Multiscale", "KuramotoSivashinskyEquation"])
def test_tensor_core_arithmetic_model_keeps_parity(harness_lib, name):
    """The numerical scheme of the tcgen05 kernels, modelled on the host: every hidden-layer contraction as bf16x3 splits, six
    products in the kernels' issue order, k-steps of 16, fp32 accumulate that TRUNCATES (mode 1) or rounds to nearest (mode 2).
    Parity against the reference's golden vectors holds under the truncating model, and the comparison of the two modes shows
    what the truncation costs (the model reproduces the GPU kernels' distance from fp64 at trained weights: DESIGN section 5)."""
    G = Golden(name)
    seeds = np.eye(G.desc.n_pde, dtype=np.float32)
    err = {}
    try:
        for mode in (1, 2):
            harness_lib.harness_set_tc_mode(mode)
            loss, grad, resid = _run(harness_lib, G, seeds)
            assert_parity(resid, G.g["resid64"], f"residual (mode {mode})")
            assert_parity(loss, G.g["loss64"], f"loss (mode {mode})")
            assert_parity(grad[:, G.idx], G.g["grad64"], f"grad (mode {mode})")
            err[mode] = float(np.linalg.norm(grad[:, G.idx] - G.g["grad64"]) / np.linalg.norm(G.g["grad64"]))
    finally:
        harness_lib.harness_set_tc_mode(0)
    assert err[2] < 1e-6 and err[1] < 1e-5
    assert err[1] > err[2]            # the truncating accumulate is the less accurate one


_TABLE = np.array([[0.0, 0.0, 1.0], [1.0, 1.0, 2.0], [0.0, 1.0, 1.5], [1.0, 0.0, 1.2]])


@pytest.mark.parametrize("H", [64, 128])
@pytest.mark.parametrize("name", ["Burgers1D", "Burgers2D", "GrayScottEquation", "Poisson1D", "PoissonND", "HeatND", "NS2D_LongTime",
                                  "Wave2D_Heterogeneous", "PoissonInv", "HeatInv", "Poisson3D_ComplexGeometry", "KuramotoSivashinskyEquation"])
def test_device_math_at_other_hidden_widths(harness_lib, name, H):
    """The jet recurrences, residual functors and adjoints are width-generic: hidden widths 64 and 128 (benchmark.py:65
    --hidden-layers) against the fp64 reference algorithm, on the host build (the GPU side: tests/test_cuda_widths.py)."""
    from golden import common
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    spec = problems.BUILDERS[name](*((_TABLE,) if name == "Wave2D_Heterogeneous" else ()))
    desc = spec.desc.with_net(H, 5)
    flat64 = common.golden_params(desc.layer_shapes(), 7 + H)
    x = spec.sample(97, seed=H)
    xt = torch.as_tensor(x)
    aux = spec.aux(xt)
    S = desc.n_pde
    ol, og, orr = O.losses_and_grads(desc, torch.as_tensor(flat64), xt.double(), None if aux is None else aux.double(), seeds=np.eye(S))
    flat = flat64.astype(np.float32)
    xx = np.ascontiguousarray(x.astype(np.float32))
    auxn = None if aux is None else np.ascontiguousarray(aux.numpy().astype(np.float32))
    n, P = xx.shape[0], desc.n_params
    grad = np.zeros((S, P)); loss = np.zeros(3); resid = np.zeros((n, S), np.float32)
    seeds = np.eye(S, dtype=np.float32)
    cd = desc.to_c()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = harness_lib.harness_run(ctypes.byref(cd), vp(flat), vp(xx), vp(auxn), ctypes.c_longlong(n), ctypes.c_double(1.0 / n), vp(seeds), S,
                                 vp(grad), vp(loss), vp(resid))
    assert rc == 0
    assert_parity(resid, orr.numpy(), "residual")
    assert_parity(loss[:S], ol.numpy(), "loss")
    assert_parity(grad, og.numpy(), "grad")
