"""Rank candidate tensor-core arithmetic schemes on the CPU: gradient error of the host build of the device math
(tests/host_harness, harness_set_tc_mode) against the fp64 reference algorithm at TRAINED weights
(gpurun_out/trained_<problem>.pt, written by scripts/accuracy_after_training.py on the GPU).

    python scripts/host_accuracy_model.py [--modes 0,1,2] [--points 2048] Poisson2D_Classic ...
"""
import argparse, ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import autograd_ref as O
from pinnacle_b200 import problems

ap = argparse.ArgumentParser()
ap.add_argument("names", nargs="*", default=["Poisson2D_Classic"])
ap.add_argument("--modes", default="0,1,2")
ap.add_argument("--points", type=int, default=2048)
ap.add_argument("--weights-dir", default=os.path.join(ROOT, "gpurun_out"))
a = ap.parse_args()
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "host_harness", "libharness.so"))
vp = lambda t: t.ctypes.data_as(ctypes.c_void_p) if t is not None else None
for name in a.names:
    spec = problems.make(name)
    d = spec.desc.with_net(100, 5)
    flat = torch.load(os.path.join(a.weights_dir, f"trained_{name}.pt")).float().cpu()
    x = spec.sample(2048, seed=5)[:a.points]
    n = x.shape[0]
    seeds = np.ones((1, d.n_pde))
    aux = spec.aux(torch.as_tensor(x)).float().numpy() if d.n_aux else None
    _, og, _ = O.losses_and_grads(d, flat.double(), torch.as_tensor(x).double(), None if aux is None else torch.as_tensor(aux).double(), seeds=seeds)
    _, og32, _ = O.losses_and_grads(d, flat, torch.as_tensor(x), None if aux is None else torch.as_tensor(aux), seeds=seeds)
    ref = og[0].numpy()
    rel = lambda g: float(np.linalg.norm(g - ref) / np.linalg.norm(ref))
    msg = f"{name:22s} |g| {np.linalg.norm(ref):.2e} ref-fp32 {rel(og32[0].double().numpy()):.2e}"
    cd = d.to_c()
    fl = np.ascontiguousarray(flat.numpy()); xx = np.ascontiguousarray(x.astype(np.float32)); sd = np.ascontiguousarray(seeds.astype(np.float32))
    for mode in a.modes.split(","):          # "m" or "fwd:bwd"
        f, b = (mode.split(":") + [mode])[:2]
        lib.harness_set_tc_modes(int(f), int(b))
        grad = np.zeros((1, d.n_params)); loss = np.zeros(3)
        rc = lib.harness_run(ctypes.byref(cd), vp(fl), vp(xx), vp(aux), ctypes.c_longlong(n), ctypes.c_double(1.0 / n), vp(sd), 1, vp(grad), vp(loss), None)
        assert rc == 0
        msg += f"  mode{mode} {rel(grad[0]):.2e}"
    print(msg, flush=True)
