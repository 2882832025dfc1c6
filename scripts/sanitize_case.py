"""Tiny invocation of both kernels for compute-sanitizer (one tool per gpurun call)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pinnacle_b200 import capi, problems
dev = torch.device("cuda:0")
for name in ("Poisson2D_Classic", "NS2D_LidDriven", "Heat2D_Multiscale", "KuramotoSivashinskyEquation", "PoissonInv"):
    spec = problems.make(name)
    d = spec.desc
    g = torch.Generator().manual_seed(0)
    flat = (torch.randn(d.n_params, generator=g) * 0.1).to(dev)
    x = torch.as_tensor(spec.sample(70, seed=1)).to(dev)
    for impl in (capi.IMPL_FFMA, capi.IMPL_TC, capi.IMPL_TC_PIPELINED):
        capi.set_impl(impl)
        out = capi.fwd_bwd(d, flat, x, None, 1.0 / 70, torch.ones(2, d.n_pde, device=dev))
        loss, res = capi.fwd(d, flat, x, None, 1.0 / 70, True)
        torch.cuda.synchronize()
        print(name, impl, float(out.abs().sum()), float(loss.sum()))

# optimizer-side kernels
P = 40801
gg = torch.randn(8, P, device=dev)
w = torch.ones(8, device=dev); w2 = torch.empty(8, device=dev); cnt = torch.full((8,), float(P), device=dev); out = torch.empty(P, device=dev)
capi.lra_combine(gg, 3, w, cnt, 0.1, w2, out)
print("stats", capi.grad_stats(out).cpu().tolist())
p_ = torch.randn(P, device=dev); m_ = torch.zeros(P, device=dev); v_ = torch.zeros(P, device=dev); st = torch.zeros(2, device=dev)
capi.adam_step(p_, out, m_, v_, st, 1e-3, 0.9, 0.999, 1e-8, 0.0)
torch.cuda.synchronize()
print("adam step", float(st[0]))
