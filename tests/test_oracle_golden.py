"""The oracle (oracle/autograd_ref.py) against vectors produced by the unmodified reference
(tests/golden/*.npz, generator tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import autograd_ref as O
from util import GOLDEN_NAMES, Golden, assert_parity, rel_l2


def test_fixture_inventory():
    assert len(GOLDEN_NAMES) == 26


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_fp64_matches_reference(name):
    G = Golden(name)
    flat, x, aux = G.inputs(torch.float64)
    losses, grads, res = O.losses_and_grads(G.desc, flat, x, aux)
    # float64 reference outputs; coefficient tables go through float32 in the reference (poisson.py:251)
    tol = 1e-6 if name in ("Poisson2D_ManyArea", "Heat2D_VaryingCoef", "Wave2D_Heterogeneous") else 1e-10
    assert rel_l2(res.numpy(), G.g["resid64"]) < tol
    assert rel_l2(losses.numpy(), G.g["loss64"]) < tol
    assert rel_l2(grads.numpy()[:, G.idx], G.g["grad64"]) < tol
    assert rel_l2(grads.numpy() @ G.probes.T, G.g["gproj64"]) < max(tol, 1e-9)
    assert rel_l2(grads.norm(dim=1).numpy(), G.g["gnorm64"]) < tol


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_fp32_within_tolerance(name):
    G = Golden(name)
    flat, x, aux = G.inputs(torch.float32)
    losses, grads, res = O.losses_and_grads(G.desc, flat, x, aux)
    assert_parity(res.numpy(), G.g["resid32"], "residual vs reference fp32")
    assert_parity(losses.numpy(), G.g["loss32"], "loss vs reference fp32")
    assert_parity(grads.numpy()[:, G.idx], G.g["grad32"], "grad vs reference fp32")
    assert_parity(res.numpy(), G.g["resid64"], "residual vs reference fp64")
    assert_parity(grads.numpy()[:, G.idx], G.g["grad64"], "grad vs reference fp64")


def test_seeded_backward_is_linear_combination():
    G = Golden("NS2D_LidDriven")
    flat, x, aux = G.inputs(torch.float64)
    _, g_eye, _ = O.losses_and_grads(G.desc, flat, x, aux)
    seeds = np.array([[0.3, -1.2, 2.0], [1.0, 1.0, 1.0]])
    _, g_s, _ = O.losses_and_grads(G.desc, flat, x, aux, seeds=seeds)
    assert rel_l2(g_s.numpy(), seeds @ g_eye.numpy()) < 1e-12
