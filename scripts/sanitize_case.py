"""Tiny invocation of both kernels for compute-sanitizer (one tool per gpurun call)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pinnacle_b200 import capi, problems
dev = torch.device("cuda:0")
for name in ("Poisson2D_Classic", "NS2D_LidDriven", "Heat2D_Multiscale"):
    spec = problems.make(name)
    d = spec.desc
    g = torch.Generator().manual_seed(0)
    flat = (torch.randn(d.n_params, generator=g) * 0.1).to(dev)
    x = torch.as_tensor(spec.sample(70, seed=1)).to(dev)
    for impl in (capi.IMPL_FFMA, capi.IMPL_TC):
        capi.set_impl(impl)
        out = capi.fwd_bwd(d, flat, x, None, 1.0 / 70, torch.ones(2, d.n_pde, device=dev))
        loss, res = capi.fwd(d, flat, x, None, 1.0 / 70, True)
        torch.cuda.synchronize()
        print(name, impl, float(out.abs().sum()), float(loss.sum()))
