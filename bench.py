#!/usr/bin/env python
"""bench.py -- train-step collocation points / second of the fused PDE-residual + gradient path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl fused|reference]

One "step" = one pass of the hot path (per-term losses + full flat parameter gradient, + the
all-reduce when N > 1) over the synthetic collocation sets of BASELINE.json configs[1]:
Poisson2D_Classic AND Heat2D_Multiscale, `--points` (default 1 Mi) uniform points each, 5x100 tanh
MLP at seeded Glorot-normal weights with N(0, 0.1^2) biases.  N > 1: one process per GPU under
torchrun, every rank holds the same per-GPU point count (weak scaling), one NCCL all-reduce of the
flat [grads | losses] buffer per problem per step.

`value`   device-resident inputs, CUDA-event time of the K steps, max over ranks.
`e2e`     the same metric through the public API (FusedPDESolver.train_step: H2D of the step's
          points from pinned host memory, fused op, autograd backward, Adam update, D2H of the losses).
`--impl reference` times the UNMODIFIED reference on the host cores: the real `dde.Model` of each workload class built
the way benchmark.py:84-122 builds it (baseline/_ref, the byte-for-byte copy scripts/install_reference.py places; imported
through baseline/loader.py) and stepped through its own `model._train_step` (deepxde/model.py:501-509), every thread torch can
use, on a bounded sample (`--ref-points` domain points per class).  It runs in a child process with CUDA hidden, because
importing the reference on a CUDA box makes CUDA the default tensor type.  Where baseline/_ref is absent the port of the same
algorithm (oracle/autograd_ref.py) is timed instead and the line says `kind: "port"`.

The default line also carries `config.extra_workloads`: BASELINE configs[4] (Poisson3D_ComplexGeometry + Heat2D_LongTime, 16 Mi
points per problem IN TOTAL, strong scaling) and configs[2] (NS2D_LidDriven + learning-rate annealing, 1 Mi residual rows in
total, strong scaling), 5 steps each with their own clock samples.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import math
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "train-step collocation pts/sec"
UNIT = "points/s"
WORKLOADS = {   # BASELINE.json configs[1] (default, the config `metric` is quoted on) and configs[4] (scaling study)
    "c2": ("Poisson2D_Classic", "Heat2D_Multiscale"),
    "c5": ("Poisson3D_ComplexGeometry", "Heat2D_LongTime"),
    "c3": ("NS2D_LidDriven",),      # + LRA, see run_c3
}
WORKLOAD_PROBLEMS = WORKLOADS["c2"]


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def _ncu_traffic(problem: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the fused kernel of `problem`, from the
    committed `ncu --set full` capture of this same command (profiles/ncu_traffic.json names the capture)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    ent = d.get("kernels", {}).get(problem)
    return (ent["dram_bytes"], d.get("source")) if ent else (None, None)


def _seeded_params(desc, seed=0):
    """Glorot-normal weights + N(0,0.1^2) biases ("trained-like" variant), flat fp32."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for (n_out, n_in) in desc.layer_shapes():
        std = (2.0 / (n_in + n_out)) ** 0.5
        parts.append((torch.randn(n_out, n_in, generator=g) * std).reshape(-1))
        parts.append(torch.randn(n_out, generator=g) * 0.1)
    return torch.cat(parts).float()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        # nvidia-smi needs ~0.1-0.3 s before its first line: wait for it, so that even a short timed region is sampled
        t0 = time.time()
        while not self.rows and time.time() - t0 < 3.0:
            time.sleep(0.01)
        self.rows.clear()                           # samples taken before the timed region started do not count

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(points_per_problem: int, steps: int, warmup: int):
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # Deep-layer jet terms go denormal on some inputs and the x86 denormal slow path costs the CPU path
    # 10-50x; flushing them is the setting most favourable to the CPU baseline (the reference leaves it off).
    torch.set_flush_denormal(True)
    specs = [problems.make(n) for n in WORKLOAD_PROBLEMS]
    work = []
    for i, spec in enumerate(specs):
        flat = _seeded_params(spec.desc, seed=i)
        x = torch.as_tensor(spec.sample(points_per_problem, seed=100 + i))
        work.append((spec.desc, flat, x))

    def step():
        for desc, flat, x in work:
            O.train_step_cpu(desc, flat, x, None)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    pts = points_per_problem * len(specs)
    return pts / dt, dt * 1e3, cores, pts


def gpu_autograd_run(points_per_problem: int, steps: int, warmup: int, dev):
    """The reference ALGORITHM (nested reverse-mode autograd, oracle/autograd_ref.py) with stock PyTorch CUDA kernels on
    the same GPU -- what a PINNacle user runs today.  Bounded sample: the retained graph is ~70 KB/point."""
    from oracle import autograd_ref as O
    from pinnacle_b200 import problems
    specs = [problems.make(n) for n in WORKLOAD_PROBLEMS]
    work = []
    for i, spec in enumerate(specs):
        work.append((spec.desc, _seeded_params(spec.desc, seed=i).to(dev), torch.as_tensor(spec.sample(points_per_problem, seed=100 + i)).to(dev)))

    def step():
        for desc, flat, x in work:
            O.train_step_cpu(desc, flat, x, None)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return points_per_problem * len(specs) / (ms * 1e-3), ms


def _reference_worker(args):
    """Child process of reference_cpu_measure(): the unmodified reference, CPU only.  Prints one `REFWORKER {json}` line."""
    from baseline import loader
    dde, _ = loader.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_flush_denormal(True)         # the setting most favourable to the CPU baseline (see cpu_reference_run)
    models, pts = [], 0
    for i, name in enumerate(WORKLOAD_PROBLEMS):
        dde.config.set_random_seed(i)
        pde = loader.construct(name)
        pde.training_points(domain=args.ref_points)                 # boundary / initial points stay at the class defaults
        net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()      # benchmark.py:94
        with torch.no_grad():
            for lin in net.linears:
                lin.bias.normal_(0, 0.1)
        with loader._in_reference_dir():
            model = pde.create_model(net)
        model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))             # benchmark.py:107,117
        net.auxiliary_vars = None
        models.append(model)
        pts += int(model.data.train_x_all.shape[0])

    def step():
        for m in models:
            m._train_step(m.data.train_x, None, None)               # deepxde/model.py:501-509

    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    print("REFWORKER " + json.dumps({"value": pts / dt, "ms": dt * 1e3, "cores": cores, "pts": pts,
                                     "rows": [int(m.data.train_x.shape[0]) for m in models]}), flush=True)


def reference_cpu_measure(args, steps: int, warmup: int):
    """The reference's CPU implementation of the step on the host cores: dict(value, ms, cores, pts, kind, sample)."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "deepxde")) and not args.port_baseline:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PINN_REFERENCE_ROOT=ref_root)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        cmd = [sys.executable, os.path.abspath(__file__), "--reference-worker", "--ref-points", str(args.ref_points), "--steps", str(steps),
               "--warmup", str(warmup), "--workload", args.workload]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
            rows = [ln for ln in r.stdout.splitlines() if ln.startswith("REFWORKER ")]
            if r.returncode == 0 and rows:
                d = json.loads(rows[-1][10:])
                d["kind"] = "reference"
                d["sample"] = (f"unmodified reference (baseline/_ref): dde.Model of {' + '.join(WORKLOAD_PROBLEMS)} built as benchmark.py:84-122, "
                               f"{args.ref_points} domain points per class + the classes' default boundary/initial points = {d['pts']} residual rows per step, "
                               f"model._train_step (losses, backward, Adam) x {warmup} warm-up + {steps} timed, {d['cores']} torch threads, denormals flushed")
                return d
            sys.stderr.write("reference worker failed, timing the port instead:\n" + r.stdout[-1000:] + r.stderr[-2000:] + "\n")
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"reference worker failed ({e!r}), timing the port instead\n")
    n = args.cpu_points
    value, ms, cores, pts = cpu_reference_run(n, steps, warmup)
    return {"value": value, "ms": ms, "cores": cores, "pts": pts, "kind": "port",
            "sample": f"port of the reference algorithm (oracle/autograd_ref.py, nested torch.autograd): {n} points of each of {'+'.join(WORKLOAD_PROBLEMS)} "
                      f"per step, {warmup} warm-up + {steps} timed steps, {cores} torch threads, denormals flushed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = reference_cpu_measure(args, max(1, args.steps), max(1, args.warmup))
    value, ms, cores = m["value"], m["ms"], m["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "+".join(WORKLOAD_PROBLEMS) + f", {m['pts']} residual rows per step (bounded CPU sample of the 1Mi-point config: the reference "
                                                             "holds the whole autograd graph in RAM; its pts/s is flat in N), 5x100 tanh MLP",
                   "device": "host CPU", "threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": m["kind"], "sample": m["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# fused arm
# ------------------------------------------------------------------------------------------------
def run_fused(args):
    import torch.distributed as dist
    from pinnacle_b200 import capi, problems
    from pinnacle_b200.fused import FusedPDEResidual
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.solver import FusedPDESolver
    from pinnacle_b200.optim import FlatAdam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (fused arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints there)
        dist.init_process_group("nccl", device_id=dev)
    capi.load()
    capi.set_impl({"auto": capi.IMPL_AUTO, "ffma": capi.IMPL_FFMA, "tc": capi.IMPL_TC, "tc2": capi.IMPL_TC_PIPELINED}[args.kernel])

    # weak scaling (default, the driver's contract): --points per problem PER GPU; strong: --points per problem in TOTAL
    n_per = args.points if args.scaling == "weak" else args.points // world
    specs = [problems.make(n) for n in WORKLOAD_PROBLEMS]
    ops, params, seeds, outs = [], [], [], []
    for i, spec in enumerate(specs):
        op = FusedPDEResidual(spec, 100, 5, device=dev)
        x_local = torch.as_tensor(spec.sample(n_per, seed=100 + i + 1000 * rank))
        op.set_points(x_local, shard=False)
        op.n_global = n_per * world
        op.inv_n_global = 1.0 / op.n_global
        ops.append(op)
        params.append(_seeded_params(spec.desc, seed=i).to(dev))
        seeds.append(torch.ones(1, spec.desc.n_pde, device=dev))
        outs.append(torch.empty(spec.desc.n_params + spec.desc.n_pde, device=dev))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    launches = [0]

    def one_step(ev=None):
        for i, op in enumerate(ops):
            if ev is not None:
                ev[i][0].record()
            capi.fwd_bwd(op.desc, params[i], op.x, op.aux, op.inv_n_global, seeds[i], out=outs[i])
            launches[0] += capi.last_launch_count()
            if ev is not None:
                ev[i][1].record()
            if world > 1:
                dist.all_reduce(outs[i])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        one_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    one_step()                                # back under load after the sampler's start-up wait (untimed)
    K = args.steps
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kern_ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ops] for _ in range(K)]
    launches[0] = 0
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()                         # L2 flush between timed iterations (outside the event pair)
        step_ev[k][0].record()
        one_step(kern_ev[k])
        step_ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = sum(a.elapsed_time(b) for a, b in step_ev)
    kern_ms = [sum(kern_ev[k][i][0].elapsed_time(kern_ev[k][i][1]) for k in range(K)) / K for i in range(len(ops))]
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    pts_step = n_per * len(ops) * world
    value = pts_step / (ms_step * 1e-3)
    gpu_launches = launches[0]

    # ---------------- e2e through the public API: host points -> losses on host ----------------
    solvers = []
    for i, spec in enumerate(specs):
        net = FNN([spec.desc.d] + [100] * 5 + [spec.desc.m]).to(dev)
        off = 0
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(params[i][off:off + p.numel()].view_as(p))
                off += p.numel()
        xh = torch.as_tensor(spec.sample(n_per, seed=100 + i + 1000 * rank)).pin_memory()
        # cache_points=False: the step's points cross PCIe every step, as in the reference (model.py:263)
        s = FusedPDESolver(spec, net, xh, cache_points=False, n_global=n_per * world)
        s.compile(FlatAdam(net.parameters(), 1e-3))              # the repo's Adam (one launch over the flat buffer, torch.optim.Adam semantics)
        solvers.append(s)

    def e2e_step2():
        for s in solvers:
            s.train_step()                                        # H2D (pinned, on the op's upload stream) + fused op + Adam
        return sum(float(s.opt.losses.detach().cpu().sum()) for s in solvers)      # D2H read of the step's results

    for _ in range(2):
        e2e_step2()
    barrier()
    Ke = max(2, min(K, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Ke):
        e2e_step2()
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item()) / Ke
    h2d = sum(n_per * s.spec.desc.d * 4 for s in solvers)
    d2h = sum(s.num_loss * 4 for s in solvers)
    e2e = {"value": pts_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "api": "FusedPDESolver.train_step with optim.FlatAdam (Adam update included), points from pinned host memory"}

    # ---------------- BASELINE configs[4] and configs[2], strong scaling, 5 steps each (every rank takes part) ----------------
    extra = None
    if not args.no_extra and args.workload == "c2" and args.scaling == "weak":
        del solvers
        torch.cuda.empty_cache()
        extra = {}
        for key, fn, n_tot in (("configs[4] strong", measure_c5, 1 << 24), ("configs[2] strong", measure_c3, 1 << 20),
                               ("configs[3] strong", measure_c4, 1 << 20), ("redraw of the residual points", measure_redraw, 1 << 24)):
            try:
                extra[key] = fn(n_tot, 5, 3, dev, world, rank, local_rank)
            except Exception as e:  # noqa: BLE001
                extra[key] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peaks, peak_src = _peaks()
    dom = int(np.argmax(kern_ms))
    desc = specs[dom].desc
    on_tc = args.kernel != "ffma" and capi.has_tensor_core_kernel(ops[dom].desc)
    flops_launch = desc.flops_step_per_point() * n_per
    ach = flops_launch / (kern_ms[dom] * 1e-3) / 1e12
    try:
        ffma_peak = capi.ffma_peak_tflops(8192)
    except Exception:  # noqa: BLE001
        ffma_peak = None
    tensor_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))

    def kinfo(i):
        d = ops[i].desc
        tc_i = args.kernel != "ffma" and capi.has_tensor_core_kernel(d)
        t = kern_ms[i] * 1e-3
        out = {"problem": specs[i].name, "kernel": capi.IMPL_NAMES.get(capi.selected_impl(d), "?"),
               "launch_ms": kern_ms[i], "achieved_tflops": d.flops_step_per_point() * n_per / t / 1e12,
               "hbm_gbs_algorithmic": d.bytes_per_point() * n_per / t / 1e9}
        out["jet_channels"] = {"canonical": d.n_chan, "propagated": capi.effective_channels(d)}
        if tc_i:
            out["executed_tensor_tflops"] = d.tc_executed_flops_per_point(1, capi.effective_channels(d)) * n_per / t / 1e12
        return out

    traffic, traffic_src = _ncu_traffic(specs[dom].name) if n_per == (1 << 20) else (None, None)
    if on_tc:
        c_eff = capi.effective_channels(ops[dom].desc)
        exe_pp = desc.tc_executed_flops_per_point(1, c_eff)
        exe = exe_pp * n_per / (kern_ms[dom] * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak if tensor_peak else None,
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_hbm_bytes": desc.bytes_per_point() * n_per,
            "kernel": f"{'pinn_tc2_kernel' if capi.selected_impl(desc) == 3 else 'pinn_tc_kernel'}<{specs[dom].name}, fwd+reverse> "
                      "(tcgen05 kind::f16, bf16x3 split precision, TMEM accumulators)",
            "launch_ms": kern_ms[dom], "algorithmic_flops_per_point": desc.flops_step_per_point(), "points_per_launch": n_per,
            "peak_source": peak_src + " bf16 dense, sustained",
            "executed_tensor_flops_per_point": exe_pp, "executed_tensor_tflops": exe,
            "jet_channels": {"canonical": desc.n_chan, "propagated": c_eff},
            "tensor_pipe_frac_executed": exe / tensor_peak if tensor_peak else None,
            "note": "algorithmic flops use the canonical per-direction jet count C = 1 + sum(order) (SURVEY 8d); the kernel propagates "
                    f"{c_eff} channels (forward-Laplacian collapse where the residual allows it) and runs each fp32-accuracy contraction as "
                    "6 bf16 MMAs per k-step on a 128x112-padded 100x100 layer: executed/algorithmic = "
                    f"{exe_pp / desc.flops_step_per_point():.2f}x",
            "ffma_peak_tflops": ffma_peak, "frac_of_ffma_peak": ach / ffma_peak if ffma_peak else None,
            "per_kernel": [kinfo(i) for i in range(len(ops))],
        }
    else:
        roofline = {
            "bound": "fp32_ffma", "achieved": ach, "peak": ffma_peak, "unit": "TFLOP/s", "frac": (ach / ffma_peak) if ffma_peak else None,
            "traffic": None,
            "kernel": f"pinn_fused_kernel<{specs[dom].name}, fwd+reverse>", "launch_ms": kern_ms[dom],
            "algorithmic_flops_per_point": desc.flops_step_per_point(), "points_per_launch": n_per,
            "peak_source": "FP32 FFMA probe kernel run in this process (pinn_ffma_peak_tflops); nominal 148*128*2*f_SM",
            "tensor_peak_bf16": tensor_peak, "frac_of_tensor_peak_bf16": ach / tensor_peak if tensor_peak else None,
            "tensor_peak_source": peak_src,
            "per_kernel": [kinfo(i) for i in range(len(ops))],
        }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        m = reference_cpu_measure(args, 3, 1)
        cpu = {"value": m["value"], "unit": UNIT, "cores": m["cores"], "kind": m["kind"], "ms_per_step": m["ms"], "sample": m["sample"]}

    gpu_ref = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, ms = gpu_autograd_run(args.gpu_ref_points, 3, 2, dev)
            gpu_ref = {"value": v, "unit": UNIT, "ms_per_step": ms, "kind": "port",
                       "sample": f"{args.gpu_ref_points} points of each problem per step; reference algorithm (nested torch.autograd) with stock "
                                 "PyTorch CUDA kernels on this GPU"}
        except Exception as e:  # noqa: BLE001
            gpu_ref = {"error": repr(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{' + '.join(WORKLOAD_PROBLEMS)}, {n_per} synthetic collocation points each per GPU, 5x100 tanh MLP, "
                               "losses + flat parameter gradient (+ NCCL all-reduce when N>1)",
                   "points_per_step": pts_step, "l2": "flushed between timed iterations (256 MiB write outside the event pair)",
                   "parallelism": f"dp{world} over collocation points", "wall_s_timed_region": t_wall, "kernel": args.kernel,
                   "extra_workloads": extra},
        "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_autograd_baseline": gpu_ref,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_c3(N, steps, warmup, dev, world, rank, local_rank):
    """BASELINE.json configs[2]: NS2D_LidDriven (u, v, p; 3 coupled residuals + 5 boundary terms incl. a one-point pressure pin)
    with learning-rate-annealing loss re-weighting, `N` residual rows IN TOTAL sharded over the ranks (strong scaling),
    through the public API: FusedPDESolver + optim.FusedLRA(FlatAdam).  One step = per-term gradient rows (one NCCL
    all-reduce of the PDE rows), the LRA statistics / re-weighting kernel, the Adam kernel.  Needs an initialised process
    group when world > 1.  Returns the record on rank 0, None elsewhere."""
    import torch.distributed as dist
    from pinnacle_b200 import problems
    from pinnacle_b200.fused import shard_bounds
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.optim import FlatAdam, FusedLRA
    from pinnacle_b200.solver import FusedPDESolver, ValueBC
    spec = problems.make("NS2D_LidDriven")
    nb = 1024
    beg, end = shard_bounds(N, rank, world)
    x_local = spec.sample(N, seed=7)[beg:end]                     # every rank draws the same set and keeps its block
    rs = np.random.RandomState(11)
    xb = rs.uniform(0, 1, (4 * nb, 2)).astype(np.float32)
    xb[:nb, 1], xb[nb:2 * nb, 1], xb[2 * nb:3 * nb, 0], xb[3 * nb:, 0] = 1.0, 0.0, 0.0, 1.0
    lid = (4.0 * xb[:nb, :1] * (1.0 - xb[:nb, :1])).astype(np.float32)       # ns.py lid profile a*x*(1-x), a = 4
    zero = lambda n: np.zeros((n, 1), np.float32)
    terms = [ValueBC(0, nb, 0, lid), ValueBC(0, nb, 1, zero(nb)), ValueBC(nb, 4 * nb, 0, zero(3 * nb)), ValueBC(nb, 4 * nb, 1, zero(3 * nb)),
             ValueBC(4 * nb - 1, 4 * nb, 2, zero(1))]
    torch.manual_seed(0)
    net = FNN([2] + [100] * 5 + [3]).to(dev)
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    w = np.ones(8)
    solver = FusedPDESolver(spec, net, x_local, xb, terms, n_global=N)
    solver.compile(FusedLRA(FlatAdam(net.parameters(), 1e-3), w, 3), loss_weights=w)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        solver.train_step()
    # The whole step -- fused passes, the NCCL all-reduce of the PDE rows, the LRA statistics and the Adam update: ~30 launches
    # around 1.3 ms of kernels at 8 GPUs -- replayed as ONE CUDA graph (FusedPDESolver.capture_step; PINN_C3_GRAPH=0: eager)
    # Single process only by default: with NCCL kernels captured in the graph the measurement itself ran (2 GPUs: same step time as
    # eager), but the processes then hung in teardown (ProcessGroupNCCL watchdog against the graph's event destruction) until killed
    # -- PINN_C3_GRAPH=force captures under torchrun anyway.
    graphed = False
    want_graph = os.environ.get("PINN_C3_GRAPH", "1")
    if want_graph == "force" or (want_graph != "0" and world == 1):
        try:
            solver.capture_step(warmup=2)
            graphed = True
        except Exception as e:  # noqa: BLE001
            graphed = repr(e)[:200]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        solver.train_step()
    e1.record()
    barrier()
    est = torch.tensor([e0.elapsed_time(e1) / 3.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    # long enough for the clock sampler (20 ms period) to see the run under load: >= 1.2 s of steps, every rank the same count
    K = int(min(max(steps, math.ceil(1200.0 / max(float(est.item()), 1e-3))), 4000))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver.train_step()                       # back under load after the sampler's start-up wait (untimed)
    barrier()
    e0.record()
    for _ in range(K):
        solver.train_step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    losses = solver.opt.losses.detach().cpu().tolist()
    weights_after = solver.opt.sync_loss_weight().tolist()
    if graphed is True:                                   # release the graph before anything else of the process goes away
        solver._graph = None
        torch.cuda.synchronize()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": N / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"NS2D_LidDriven + learning-rate annealing (FusedLRA over FlatAdam), {N} residual rows in total + {4 * nb} boundary "
                               "rows (replicated), 3 PDE + 5 boundary terms, 5x100 tanh MLP; optimizer update included",
                   "parallelism": f"dp{world} over collocation points", "l2": "inputs larger than L2 at 1 GPU; no flush between steps",
                   "cuda_graph": graphed},
        "clocks": clocks, "loss_weights_after": weights_after, "losses_after": losses}


def measure_redraw(total_points, steps, warmup, dev, world, rank, local_rank):
    """PDEPointResampler's redraw (deepxde/callbacks.py:563-568) on the device: `total_points` rows x 3 of Heat2D_ComplexGeometry's
    domain (rectangle minus 17 disks, times the time interval) and of its plain bounding box, split over the ranks by row range (a
    row is a function of its global index: no exchange), one pinn_sample_points launch per rank, in place.  CUDA events, max over
    ranks."""
    import torch.distributed as dist
    from pinnacle_b200 import capi
    from pinnacle_b200.fused import shard_bounds
    from pinnacle_b200.resample import Region
    lo, hi = (-8.0, -12.0, 0.0), (8.0, 12.0, 3.0)
    disks = [(0, (cx, cy), (1.0,)) for cx, cy in ((-4, -3), (4, -3), (-4, 3), (4, 3), (-4, -9), (4, -9), (-4, 9), (4, 9), (0, 0), (0, 6), (0, -6))]
    disks += [(0, (cx, cy), (0.4,)) for cx, cy in ((-3.2, -6), (-3.2, 6), (3.2, -6), (3.2, 6), (-3.2, 0), (3.2, 0))]
    beg, end = shard_bounds(total_points, rank, world)
    x = torch.empty((end - beg, 3), dtype=torch.float32, device=dev)
    out = {"rows": total_points, "bytes_written": total_points * 12}
    for tag, reg in (("box_minus_17_disks_ms", Region(lo, hi, disks)), ("plain_box_ms", Region(lo, hi))):
        c = reg.to_c()
        for i in range(max(warmup, 3)):
            capi.sample_points(x, c, 7, i, beg)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(steps):
            capi.sample_points(x, c, 7, 100 + i, beg)
        ev[1].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[tag] = float(t.item())
    out["plain_box_gbs_written"] = out["bytes_written"] / (out["plain_box_ms"] * 1e-3) / 1e9
    out["note"] = "the reference's host redraw of the same 16 Mi rows: 0.69 s (box) / 17.9 s (17 disks) of numpy + a 201 MB upload (profiles/r2_resample.md)"
    return out


def measure_c4(total_points, steps, warmup, dev, world, rank, local_rank):
    """BASELINE.json configs[3]: Kuramoto-Sivashinsky (fourth-order jets in x), `total_points` synthetic collocation points IN
    TOTAL split over the ranks, Adam then L-BFGS through the public API: FusedPDESolver + optim.FlatAdam for the Adam phase, then
    torch.optim.LBFGS (lr 1, max_iter 20, no line search: what benchmark.py:114-115 / src/optimizer/adam_lbfgs.py hand to the
    reference) driving the fused closure.  Reported: ms per Adam step (the line's ms_per_step / value) and ms per L-BFGS closure
    evaluation (one fused forward + reverse + all-reduce, the two-loop recursion and the update of torch's own L-BFGS)."""
    import torch.distributed as dist
    from pinnacle_b200 import problems
    from pinnacle_b200.fused import shard_bounds
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.optim import FlatAdam
    from pinnacle_b200.solver import FusedPDESolver
    spec = problems.make("KuramotoSivashinskyEquation")
    beg, end = shard_bounds(total_points, rank, world)
    x_local = spec.sample(total_points, seed=9)[beg:end]
    torch.manual_seed(0)
    net = FNN([2] + [100] * 5 + [1]).to(dev)
    with torch.no_grad():
        for lin in net.linears:
            lin.bias.normal_(0, 0.1)
    solver = FusedPDESolver(spec, net, x_local, n_global=total_points).compile(FlatAdam(net.parameters(), 1e-3))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        solver.train_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver.train_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = max(steps, 20)
    e0.record()
    for _ in range(K):
        solver.train_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    # L-BFGS phase: same solver, the reference's optimizer settings, the fused losses through autograd (closure protocol)
    solver.compile(torch.optim.LBFGS(net.parameters(), lr=1, max_iter=20))
    evals = [0]
    orig = solver.outputs_losses_train

    def counted(*a, **k):
        evals[0] += 1
        return orig(*a, **k)

    solver.outputs_losses_train = counted
    solver.train_step()                       # untimed first iteration (history empty)
    evals[0] = 0
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(3):
        solver.train_step()                   # one L-BFGS step = up to 20 closure evaluations
    e3.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t2 = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    n_ev = max(evals[0], 1)
    loss = float(solver.opt.losses.detach().sum().cpu()) if hasattr(solver.opt, "losses") else None
    if loss is not None and not math.isfinite(loss):
        loss = None                           # (an L-BFGS without line search may leave the basin; the timing stands)
    if rank != 0:
        return None
    return {"metric": METRIC, "value": total_points / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(warmup, 3),
            "ms_per_step": ms, "scaling": "strong",
            "config": {"workload": f"KuramotoSivashinskyEquation (jet orders (4,1)), {total_points} synthetic collocation points in total, 5x100 tanh MLP, "
                                   "Adam (FlatAdam, update included) then L-BFGS (torch.optim.LBFGS lr 1 max_iter 20, fused closure)",
                       "parallelism": f"dp{world} over collocation points"},
            "lbfgs": {"closure_evaluations": n_ev, "ms_per_closure_evaluation": float(t2.item()) / n_ev, "loss_after": loss},
            "clocks": clocks}


def measure_c5(total_points, steps, warmup, dev, world, rank, local_rank):
    """BASELINE.json configs[4]: Poisson3D_ComplexGeometry + Heat2D_LongTime, `total_points` collocation points per problem IN
    TOTAL split over the ranks (strong scaling): losses + flat gradient per problem (+ the all-reduce), device-resident points."""
    import torch.distributed as dist
    from pinnacle_b200 import capi, problems
    from pinnacle_b200.fused import FusedPDEResidual
    names = WORKLOADS["c5"]
    n_per = total_points // world
    work = []
    for i, name in enumerate(names):
        spec = problems.make(name)
        op = FusedPDEResidual(spec, 100, 5, device=dev)
        op.set_points(torch.as_tensor(spec.sample(n_per, seed=300 + i + 1000 * rank)), shard=False)
        op.n_global = n_per * world
        op.inv_n_global = 1.0 / op.n_global
        work.append((op, _seeded_params(spec.desc, seed=i).to(dev), torch.ones(1, spec.desc.n_pde, device=dev),
                     torch.empty(spec.desc.n_params + spec.desc.n_pde, device=dev)))

    def step():
        for op, prm, sd, out in work:
            capi.fwd_bwd(op.desc, prm, op.x, op.aux, op.inv_n_global, sd, out=out)
            if world > 1:
                dist.all_reduce(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                       # blocks until nvidia-smi delivers its first line: the other ranks must not start timing yet
    step()                                    # back under load after that wait (untimed)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    del work
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"metric": METRIC, "value": n_per * world * len(names) / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms, "scaling": "strong",
            "config": {"workload": f"{' + '.join(names)}, {n_per * world} synthetic collocation points each in total ({n_per} per GPU), 5x100 tanh MLP, losses + "
                                   "flat parameter gradient (+ NCCL all-reduce when N>1)", "l2": "inputs (2 x 200 MB in total) larger than L2; no flush between steps"},
            "clocks": clocks}


def run_c3(args):
    """`--workload c3` as a line of its own (not the driver's default line)."""
    import torch.distributed as dist
    from pinnacle_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    capi.load()
    rec = measure_c3(args.points, args.steps, args.warmup, dev, world, rank, local_rank)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=1 << 20, help="collocation points per problem per GPU")
    ap.add_argument("--cpu-points", type=int, default=16384, help="points per problem of the bounded CPU sample (port of the reference algorithm)")
    ap.add_argument("--ref-points", type=int, default=65536, help="domain points per class of the bounded CPU sample (unmodified reference)")
    ap.add_argument("--port-baseline", action="store_true", help="time the port of the reference algorithm even where baseline/_ref is installed")
    ap.add_argument("--reference-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-extra", action="store_true", help="skip config.extra_workloads (configs[4] at 16 Mi points and configs[2], strong scaling)")
    ap.add_argument("--impl", choices=["fused", "reference"], default="fused")
    ap.add_argument("--kernel", choices=["auto", "ffma", "tc", "tc2"], default="auto",
                    help="fused arm only: tcgen05 tensor-core kernel (auto/tc) or the FP32 FFMA kernel")
    ap.add_argument("--gpu-ref-points", type=int, default=131072, help="points per problem of the stock-PyTorch-on-GPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU and stock-PyTorch-GPU baselines")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2", help="c2 = BASELINE configs[1] (default); c5 = configs[4]; c3 = configs[2] (NS2D_LidDriven + LRA, strong scaling)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak", help="strong: --points is the TOTAL per problem, split over ranks")
    args = ap.parse_args()
    global WORKLOAD_PROBLEMS
    WORKLOAD_PROBLEMS = WORKLOADS[args.workload]
    if args.reference_worker:
        _reference_worker(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "c3":
        run_c3(args)
    else:
        run_fused(args)


if __name__ == "__main__":
    main()
