"""Subprocess side of tests/test_cuda_reference_model.py and of the real-reference L2RE replay (tests/test_cuda_l2re.py).

Importing the reference on a CUDA box makes CUDA the process-wide default tensor type (deepxde/backend/pytorch/tensor.py:18-19,
as trainer.py:39-41 does for benchmark.py), so everything that builds a real ``dde.Model`` on the GPU runs in a process of
its own and reports one JSON line on stdout.

    python tests/ref_gpu_driver.py train NAME METHOD STEPS [--fuse-optimizer]
    python tests/ref_gpu_driver.py l2re NAME

The reference comes from baseline/_ref (scripts/install_reference.py) -- the unmodified tree, imported through tests/refshim.py.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def build(name, method, seed=0, points=None):
    """benchmark.py:84-122 -- pde, FNN, optimizer of ``--method``, create_model, compile -- with every tensor on CUDA."""
    import torch
    import refshim
    dde, _ = refshim.load()
    from src.optimizer import MultiAdam, LR_Adaptor, LR_Adaptor_NTK, Adam_LBFGS
    dde.config.set_random_seed(seed)
    pde = refshim.construct(name)
    if points is not None:
        pde.training_points(**points)
    if method == "gepinn":                     # benchmark.py:91-92
        pde.use_gepinn()
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    with torch.no_grad():
        for lin in net.linears:                # zero biases: BC gradient mean exactly 0 -> inf in LRA
            lin.bias.copy_((torch.randn(lin.bias.shape, generator=g, device="cpu") * 0.1).to(lin.bias.device))
    loss_weights = np.ones(pde.num_loss)
    opt = torch.optim.Adam(net.parameters(), 1e-3)
    if method == "multiadam":
        opt = MultiAdam(net.parameters(), lr=1e-3, betas=(0.99, 0.99), loss_group_idx=[pde.num_pde])
    elif method == "lra":
        opt = LR_Adaptor(opt, loss_weights, pde.num_pde)
    elif method == "ntk":
        opt = LR_Adaptor_NTK(opt, loss_weights, pde)
    elif method == "lbfgs":
        opt = Adam_LBFGS(net.parameters(), switch_epoch=25, adam_param={"lr": 1e-3})
    with refshim._in_reference_dir():
        model = pde.create_model(net)
    model.compile(opt, loss_weights=loss_weights)
    return model, net, loss_weights


def flat(net):
    import torch
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).double().cpu().numpy()


def cmd_train(a):
    import torch
    from pinnacle_b200 import capi
    from pinnacle_b200.solver import enable_fused
    out = {}
    runs = {}
    for tag in ("reference", "fused"):
        model, net, w = build(a.name, a.method)
        assert next(net.parameters()).is_cuda, "the reference did not put the net on CUDA"
        if tag == "fused":
            enable_fused(model, fuse_optimizer=a.fuse_optimizer)
            out["optimizer"] = type(model.opt).__name__
            out["impl"] = capi.selected_impl(model.fused_core.op.desc)
        with tempfile.TemporaryDirectory() as tmp:
            model.train(iterations=a.steps, display_every=max(1, a.steps // 5), model_save_path=tmp, save_model=False)
        if hasattr(model.opt, "sync_loss_weight"):
            model.opt.sync_loss_weight()
        hist = np.asarray(model.losshistory.loss_train, dtype=np.float64)
        runs[tag] = dict(hist=hist, flat=flat(net), w=np.asarray(w, dtype=np.float64).copy(),
                         x=np.asarray(model.data.train_x, dtype=np.float64))
        if tag == "fused":
            st = model.fused_core.op.stats
            out["fused_calls"] = int(st["fwd_bwd_calls"] + st["fwd_calls"])
    r, f = runs["reference"], runs["fused"]
    out["same_points"] = bool(np.array_equal(r["x"], f["x"]))
    out["hist_reference"] = r["hist"].tolist()
    out["hist_fused"] = f["hist"].tolist()
    out["first_rel"] = float(np.max(np.abs(f["hist"][0] - r["hist"][0]) / np.maximum(np.abs(r["hist"][0]), 1e-30)))
    out["hist_rel"] = float(np.max(np.abs(f["hist"] - r["hist"]) / np.maximum(np.abs(r["hist"]), 1e-12)))
    out["weights_rel"] = float(np.linalg.norm(f["flat"] - r["flat"]) / np.linalg.norm(r["flat"]))
    out["loss_weights_reference"] = r["w"].tolist()
    out["loss_weights_fused"] = f["w"].tolist()
    out["finite"] = bool(np.all(np.isfinite(r["hist"])) and np.all(np.isfinite(f["hist"])))
    print("RESULT " + json.dumps(out))


def cmd_l2re(a):
    import make_l2re_golden as mk
    from pinnacle_b200.solver import enable_fused
    g = np.load(os.path.join(HERE, "golden", f"l2re_{a.name}.npz"))
    dde, pde, net, model = mk.build(a.name, g["flat_init"])
    assert next(net.parameters()).is_cuda
    same = bool(np.array_equal(np.asarray(model.data.train_x, dtype=np.float32), g["train_x"]))
    enable_fused(model)
    from src.utils.callbacks import TesterCallback
    with tempfile.TemporaryDirectory() as tmp:
        model.train(iterations=int(g["steps"]), display_every=int(g["log_every"]),
                    callbacks=[TesterCallback(log_every=int(g["log_every"]), verbose=False, fRMSE_param={"enable": False})],
                    model_save_path=tmp, save_model=False)
        err = np.loadtxt(os.path.join(tmp, "errors.txt")).reshape(-1, 10)
    st = model.fused_core.op.stats
    print("RESULT " + json.dumps(dict(same_points=same, l2re=err[:, 5].tolist(), fused_calls=int(st["fwd_bwd_calls"]))))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    t = sub.add_parser("train")
    t.add_argument("name"); t.add_argument("method"); t.add_argument("steps", type=int)
    t.add_argument("--fuse-optimizer", action="store_true")
    l = sub.add_parser("l2re")
    l.add_argument("name")
    a = ap.parse_args()
    {"train": cmd_train, "l2re": cmd_l2re}[a.cmd](a)
