"""Build libpinn_b200.so in-tree with nvcc for sm_100a (no torch dependency, plain C ABI).

    python pinnacle_b200/csrc/build.py [--force] [--jobs N] [--verbose]

One translation unit per kernel configuration (pinn_configs.h) so the 8 host cores compile in
parallel; objects are cached by source mtime under csrc/build/.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# what-if builds (scripts/kernel_time.py --lib ...): PINN_BUILD_DIR / PINN_LIB_OUT redirect the objects and the library,
# PINN_EXTRA_NVCC_FLAGS adds -D switches; the product build uses none of them
BUILD = os.environ.get("PINN_BUILD_DIR") or os.path.join(HERE, "build")
LIB = os.environ.get("PINN_LIB_OUT") or os.path.join(HERE, "libpinn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("PINN_EXTRA_NVCC_FLAGS", "").split()
HEADERS = ["jet_math.cuh", "pinn_kernels.cuh", "pinn_configs.h", "pinn_dispatch.h", "pinn_inst.inl", "pinn_helpers.cuh", "umma.cuh", "pinn_tc_kernels.cuh", "pinn_tc2_kernels.cuh",
           os.path.join("..", "..", "include", "pinn_b200.h")]


def parse_configs():
    txt = open(os.path.join(HERE, "pinn_configs.h")).read()
    out = []
    for m in re.finditer(r"^\s*X\(([^)]*)\)", txt, flags=re.M):
        args = [a.strip() for a in m.group(1).split(",")]
        if args[0] == "ID":
            continue
        out.append([int(a) for a in args])
    return out


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, verbose):
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True, cwd=HERE)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log[-6000:]}")
    if verbose:
        print(log)
    return obj


def build(force=False, jobs=None, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    deps = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    units = []
    for cfg in parse_configs():
        cid = cfg[0]
        src = os.path.join(BUILD, f"inst_{cid}.cu")
        body = ('#include "%s"\nnamespace pinn {\ntemplate <>\nconst ConfigInfo* get_config<%d>() { return Inst<%s>::info(); }\n}\n'
                % (os.path.join(HERE, "pinn_inst.inl"), cid, ", ".join(str(a) for a in cfg)))
        if not os.path.exists(src) or open(src).read() != body:
            with open(src, "w") as f:
                f.write(body)
        units.append((src, os.path.join(BUILD, f"inst_{cid}.o")))
    units.append((os.path.join(HERE, "pinn_api.cu"), os.path.join(BUILD, "pinn_api.o")))
    units.append((os.path.join(HERE, "pinn_umma_test.cu"), os.path.join(BUILD, "pinn_umma_test.o")))
    units.append((os.path.join(HERE, "pinn_optim.cu"), os.path.join(BUILD, "pinn_optim.o")))
    units.append((os.path.join(HERE, "pinn_sample.cu"), os.path.join(BUILD, "pinn_sample.o")))
    todo = [(s, o) for s, o in units if force or _newer(o, deps + [s])]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count()) as ex:
            list(ex.map(lambda so: _compile(so[0], so[1], verbose), todo))
    objs = [o for _, o in units]
    if todo or _newer(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose))
