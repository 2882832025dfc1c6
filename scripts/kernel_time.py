"""Device time of one pinn_fwd / pinn_fwd_bwd call per kernel implementation (CUDA events, L2 flushed between calls).

    python scripts/kernel_time.py [--problems Poisson2D_Classic,NS2D_LidDriven] [--impls tc,tc2] [--points 1048576] [--check]
"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pinnacle_b200 import capi, problems

ap = argparse.ArgumentParser()
ap.add_argument("--problems", default="Poisson2D_Classic,NS2D_LidDriven,Burgers1D,Heat2D_Multiscale")
ap.add_argument("--impls", default="tc,tc2")
ap.add_argument("--points", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--lib", default=None, help="alternative libpinn_b200.so (what-if builds)")
ap.add_argument("--modes", default="fwd_bwd,fwd")
ap.add_argument("--check", action="store_true", help="compare each impl's output with the first impl's")
args = ap.parse_args()
if args.lib:
    capi.LIB_PATH = os.path.abspath(args.lib)
dev = torch.device("cuda:0")
IM = {"auto": 0, "ffma": 1, "tc": 2, "tc2": 3}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in args.problems.split(","):
    gep = name.endswith("_gepinn")            # e.g. Poisson2D_Classic_gepinn: the gPINN form (problems.gpinn), embedded rows (x0, x1, 0, 0)
    base = name[:-7] if gep else name
    spec = problems.make(base, **({"darcy_2d_coef": np.random.RandomState(0).rand(64, 3)} if base == "Wave2D_Heterogeneous" else {}))
    if gep:
        spec = problems.gpinn(spec)
    d = spec.desc
    x = torch.as_tensor(spec.sample(args.points, seed=1))
    if gep:
        x = torch.cat([x, torch.zeros(x.shape[0], d.d - x.shape[1])], dim=1)
    x = x.to(dev).contiguous()
    aux = spec.aux(x)
    g = torch.Generator().manual_seed(0)
    params = (torch.randn(d.n_params, generator=g) * 0.1).to(dev)
    seeds = torch.ones(1, d.n_pde, device=dev)
    base = None
    for impl in args.impls.split(","):
        capi.set_impl(IM[impl])
        for mode in args.modes.split(","):
            def call():
                if mode == "fwd":
                    return capi.fwd(d, params, x, aux, 1.0 / args.points)[0]
                return capi.fwd_bwd(d, params, x, aux, 1.0 / args.points, seeds)
            for _ in range(2):
                out = call()
            ts = []
            for _ in range(args.iters):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = call(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            msg = f"{name:22s} {impl:5s} {mode:8s} {np.median(ts):8.3f} ms  ({args.points / np.median(ts) * 1e-3:.1f} Mpts/s)"
            if args.check and mode == "fwd_bwd":
                if base is None:
                    base = out.clone()
                else:
                    msg += f"  rel diff vs first impl {float((out - base).norm() / base.norm()):.2e}"
            print(msg, flush=True)
capi.set_impl(0)
