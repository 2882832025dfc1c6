"""Step latency at the reference's own problem sizes (BASELINE.json configs[0]: Burgers1D, ~12k residual rows + 4k BC/IC rows, Adam)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pinnacle_b200 import problems
from pinnacle_b200.nn import FNN
from pinnacle_b200.solver import FusedPDESolver, ValueBC
from oracle import autograd_ref as O
dev = torch.device("cuda:0")
for name, n_pde, n_bc in (("Burgers1D", 12288, 4096), ("Poisson2D_Classic", 10240, 2048), ("NS2D_LidDriven", 10240, 4097), ("Heat2D_Multiscale", 49152, 16385)):
    spec = problems.make(name)
    d = spec.desc
    torch.manual_seed(0)
    net = FNN([d.d] + [100] * 5 + [d.m]).to(dev)
    x = spec.sample(n_pde, seed=1)
    xb = spec.sample(n_bc, seed=2)
    tgt = torch.zeros(n_bc, 1, device=dev)
    solver = FusedPDESolver(spec, net, x, xb, [ValueBC(0, n_bc, 0, np.zeros((n_bc, 1), np.float32))]).compile(torch.optim.Adam(net.parameters(), 1e-3))
    for _ in range(20):
        solver.train_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 200
    for _ in range(K):
        solver.train_step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / K * 1e3
    # stock PyTorch GPU path of the reference algorithm, same sizes
    net2 = FNN([d.d] + [100] * 5 + [d.m]).to(dev)
    opt2 = torch.optim.Adam(net2.parameters(), 1e-3)
    xt, xbt = torch.as_tensor(x).to(dev), torch.as_tensor(xb).to(dev)
    def ref_step():
        opt2.zero_grad()
        flat = torch.cat([p.reshape(-1) for p in net2.parameters()])
        losses, _, _ = O.pde_losses(d, flat, xt, None)
        bc = torch.mean(torch.square(net2(xbt)[:, 0:1] - tgt))
        (losses.sum() + bc).backward()
        opt2.step()
    for _ in range(5):
        ref_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30):
        ref_step()
    torch.cuda.synchronize()
    ms2 = (time.perf_counter() - t0) / 30 * 1e3
    # same step replayed from a CUDA graph
    torch.manual_seed(0)
    net3 = FNN([d.d] + [100] * 5 + [d.m]).to(dev)
    s3 = FusedPDESolver(spec, net3, x, xb, [ValueBC(0, n_bc, 0, np.zeros((n_bc, 1), np.float32))]).compile(torch.optim.Adam(net3.parameters(), 1e-3, capturable=True))
    s3.capture_step()
    for _ in range(20):
        s3.train_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        s3.train_step()
    torch.cuda.synchronize()
    ms3 = (time.perf_counter() - t0) / K * 1e3
    print(f"{name:22s} graph replay {ms3:.3f} ms/step ({(n_pde+n_bc)/ms3*1e3:.3g} pts/s), loss {float(s3.opt.losses.sum()):.4g}")
    print(f"{name:22s} rows {n_pde}+{n_bc}: fused {ms:.3f} ms/step ({(n_pde+n_bc)/ms*1e3:.3g} pts/s) | stock torch GPU {ms2:.2f} ms/step | x{ms2/ms:.1f}")
    # autograd-free step with the flat-buffer optimizer (pinnacle_b200.optim.FlatAdam), eager and replayed
    from pinnacle_b200.optim import FlatAdam
    for tag, graph in (("FlatAdam eager", False), ("FlatAdam graph", True)):
        torch.manual_seed(0)
        net4 = FNN([d.d] + [100] * 5 + [d.m]).to(dev)
        s4 = FusedPDESolver(spec, net4, x, xb, [ValueBC(0, n_bc, 0, np.zeros((n_bc, 1), np.float32))]).compile(FlatAdam(net4.parameters(), 1e-3))
        if graph:
            s4.capture_step()
        for _ in range(20):
            s4.train_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            s4.train_step()
        torch.cuda.synchronize()
        ms4 = (time.perf_counter() - t0) / K * 1e3
        print(f"{name:22s} {tag}: {ms4:.3f} ms/step ({(n_pde+n_bc)/ms4*1e3:.3g} pts/s)")

# learning-rate annealing on NS2D_LidDriven (BASELINE configs[2]) at the reference's size: reference algorithm through autograd
# (one backward per term + .item() syncs) on the fused losses vs the row-based FusedLRA
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_cuda_optim import _LRAdaptorPort, _ns_problem
from pinnacle_b200.optim import FlatAdam, FusedLRA
spec, x, xb, terms = _ns_problem(dev, n=10240, nb=1024)
for tag in ("LRA via autograd (reference algorithm, fused losses)", "FusedLRA eager", "FusedLRA graph"):
    torch.manual_seed(0)
    net = FNN([2] + [100] * 5 + [3]).to(dev)
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    w = np.ones(8)
    s = FusedPDESolver(spec, net, x, xb, terms)
    if tag.startswith("LRA via"):
        s.compile(_LRAdaptorPort(torch.optim.Adam(net.parameters(), 1e-3), w, 3), loss_weights=w)
    else:
        s.compile(FusedLRA(FlatAdam(net.parameters(), 1e-3), w, 3), loss_weights=w)
        if tag.endswith("graph"):
            s.capture_step()
    for _ in range(10):
        s.train_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        s.train_step()
    torch.cuda.synchronize()
    print(f"NS2D_LidDriven 10240+4096 rows, 3 PDE + 5 BC terms, {tag}: {(time.perf_counter() - t0) / 100 * 1e3:.3f} ms/step")
