"""Gradient error vs the fp64 reference algorithm at TRAINED weights (1000 fused Adam steps), per kernel implementation,
next to the reference algorithm's own fp32 error (parity at init and after ~1k steps)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import autograd_ref as O
from pinnacle_b200 import capi, problems
from pinnacle_b200.nn import FNN
from pinnacle_b200.optim import FlatAdam
from pinnacle_b200.solver import FusedPDESolver, ValueBC
dev = torch.device("cuda:0")
if os.environ.get("PINN_LIB"):                       # what-if builds of the library
    capi.LIB_PATH = os.path.abspath(os.environ["PINN_LIB"])
for name in sys.argv[1:] or ["Poisson2D_Classic", "Burgers1D", "Heat2D_Multiscale", "NS2D_LidDriven"]:
    spec = problems.make(name)
    d = spec.desc
    torch.manual_seed(41)
    net = FNN([d.d] + [100] * 5 + [d.m])
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)           # on the CPU generator, like tests/test_cuda_solver.py::_net: the test's own end state
    net = net.to(dev)
    x = spec.sample(2048, seed=5); xb = spec.sample(256, seed=6)
    tgt = np.sin(3.0 * xb[:, :1]).astype(np.float32)
    wfile = os.path.join(ROOT, "gpurun_out", f"trained_{name}.pt")
    # committed weights of such runs: tests/golden/trained (round 2, first kernel) and tests/golden/trained_b (the end state of
    # tests/test_cuda_solver.py::test_parity_holds_after_a_thousand_adam_steps with the plain Laplacian tanh rule -- the state on
    # which that rule cost 1e-5); PINN_TRAINED_DIR selects
    fixed = os.path.join(ROOT, "tests", "golden", os.environ.get("PINN_TRAINED_DIR", "trained"), f"trained_{name}.pt")
    if os.environ.get("PINN_LOAD_TRAINED") and not os.path.exists(wfile) and os.path.exists(fixed):
        wfile = fixed
    if os.environ.get("PINN_LOAD_TRAINED") and os.path.exists(wfile):      # evaluate a what-if build at the SAME trained weights
        flat = torch.load(wfile).to(dev)
    else:
        s = FusedPDESolver(spec, net, x, xb, [ValueBC(0, 256, 0, tgt)]).compile(FlatAdam(net.parameters(), lr=2e-3))
        for _ in range(1000):
            s.train_step()
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        os.makedirs(os.path.dirname(wfile), exist_ok=True)
        torch.save(flat.cpu(), wfile)
    seeds = np.ones((1, d.n_pde))
    _, og, _ = O.losses_and_grads(d, flat.cpu().double(), torch.as_tensor(x).double(), None, seeds=seeds)
    _, og32, _ = O.losses_and_grads(d, flat.cpu().float(), torch.as_tensor(x), None, seeds=seeds)
    ref = og[0]
    rel = lambda g: float((g.double().cpu() - ref).norm() / ref.norm())
    msg = f"{name:22s} |g| {float(ref.norm()):.2e}  ref-fp32 {rel(og32[0]):.2e}"
    xt = torch.as_tensor(x).to(dev)
    sd = torch.ones(1, d.n_pde, device=dev)
    for tag, impl, lap in (("ffma", 1, True), ("ffma-nolap", 1, False), ("tc", 2, True), ("tc2", 3, True), ("tc2-nolap", 3, False)):
        capi.set_impl(impl); capi.set_laplacian_collapse(lap)
        out = capi.fwd_bwd(d.with_net(100, 5), flat, xt, None, 1.0 / x.shape[0], sd)
        msg += f"  {tag} {rel(out[:d.n_params]):.2e}"
    capi.set_impl(0); capi.set_laplacian_collapse(True)
    print(msg, flush=True)
