"""Accuracy of both kernels vs the reference's fp64 golden vectors (run on a B200)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import GOLDEN_NAMES, Golden, rel_l2, rel_max
from pinnacle_b200 import capi
dev = torch.device("cuda:0")
print(f"{'class':28s} impl   resid-L2   loss-rel   grad-L2    grad-max  | reference fp32: resid-L2 grad-L2")
for name in GOLDEN_NAMES:
    G = Golden(name); desc = G.desc
    flat, x, aux = [None if t is None else t.to(dev).contiguous() for t in G.inputs(torch.float32)]
    S, P = desc.n_pde, desc.n_params
    for impl, code in (("ffma", 1), ("tc", 2)):
        if code == 2 and not capi.has_tensor_core_kernel(desc):
            continue
        capi.set_impl(code)
        out = capi.fwd_bwd(desc, flat, x, aux, 1.0 / x.shape[0], torch.eye(S, device=dev))
        _, resid = capi.fwd(desc, flat, x, aux, 1.0 / x.shape[0], True)
        g = out[:S * P].view(S, P).cpu().numpy()[:, G.idx]
        l = out[S * P:].cpu().numpy()
        print(f"{name:28s} {impl:5s} {rel_l2(resid.cpu().numpy(), G.g['resid64']):.2e}  {rel_max(l, G.g['loss64']):.2e}  "
              f"{rel_l2(g, G.g['grad64']):.2e}  {rel_max(g, G.g['grad64']):.2e}  | {rel_l2(G.g['resid32'], G.g['resid64']):.2e} {rel_l2(G.g['grad32'], G.g['grad64']):.2e}")
capi.set_impl(0)
