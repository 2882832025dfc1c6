"""Pin the oracle to the UNMODIFIED reference, live (only where /root/reference is mounted): same weights, same
points -> residual matrix, per-term losses and per-term flat gradients of the reference pipeline
(net forward, the class's own pde() closure with dde.grad, mean-square per term, backward) vs oracle/autograd_ref.py."""
import numpy as np
import pytest
import torch

import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")

CLASSES = ["Burgers1D", "Poisson2D_Classic", "PoissonBoltzmann2D", "Poisson3D_ComplexGeometry", "Heat2D_Multiscale",
           "Heat2D_LongTime", "NS2D_LidDriven", "NS2D_LongTime", "KuramotoSivashinskyEquation", "GrayScottEquation",
           "Burgers2D", "Poisson2D_ManyArea", "Heat2D_VaryingCoef", "Wave2D_Heterogeneous", "HeatND", "Helmholtz2D",
           "PoissonInv", "HeatInv", "NS2D_Classic_linear"]


@pytest.mark.parametrize("name", CLASSES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_equals_reference(name, dtype):
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    dde, _ = refshim.load()
    pde = refshim.construct(name)
    spec = problems.describe_reference(pde)
    desc = spec.desc
    torch.manual_seed(1)
    torch.set_default_dtype(dtype)
    dde.config.set_default_float("float32" if dtype == torch.float32 else "float64")
    try:
        net = dde.nn.FNN([desc.d] + [100] * 5 + [desc.m], "tanh", "Glorot normal")
        with torch.no_grad():
            for lin in net.linears:
                lin.bias.normal_(0, 0.1)
        x = spec.sample(48, seed=2, dtype=np.float32 if dtype == torch.float32 else np.float64)
        rl, rg, rr = refshim.reference_losses_and_grads(pde, net, x)
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        ol, og, orr = O.losses_and_grads(desc, flat, torch.as_tensor(x), spec.aux(torch.as_tensor(x)))
    finally:
        torch.set_default_dtype(torch.float32)
        dde.config.set_default_float("float32")
    tol = 1e-6 if dtype == torch.float32 else (1e-7 if desc.n_aux and name in ("Poisson2D_ManyArea", "Heat2D_VaryingCoef", "Wave2D_Heterogeneous") else 1e-12)
    assert float((orr - rr).norm() / rr.norm()) < tol
    assert float(((ol - rl).abs() / rl.abs()).max()) < tol
    assert float((og - rg).norm() / rg.norm()) < tol
