"""Render the event timeline printed by a -DPINN_TRACE build of the pipelined tensor-core kernel.

    PINN_EXTRA_NVCC_FLAGS=-DPINN_TRACE python pinnacle_b200/csrc/build.py --force      # debug build (restore afterwards!)
    python scripts/kernel_time.py --problems Poisson2D_Classic --impls tc2 --iters 1 --modes fwd_bwd > trace.txt
    python scripts/trace_timeline.py trace.txt [launches_in_file=3] [sm_mhz=1965]

Columns: element-wise warp 0 | producer warp 16 | issuer warp 20; time in microseconds from the first event.
"""
import sys

NAMES = {1: "L1 A start", 2: "L1 B start", 3: "L1 A done", 4: "L1 B done", 10: "fwd wait A", 11: "fwd wait B", 12: "fwd got A", 13: "fwd got B",
         14: "fwd done A", 15: "fwd done B", 20: "rev start A", 21: "rev start B", 22: "rev arrays free A", 23: "rev arrays free B",
         24: "rev part1 done A", 25: "rev part1 done B", 26: "rev got A", 27: "rev got B", 28: "rev done A", 29: "rev done B",
         40: "P prefetch issued", 41: "P operand free", 42: "P operand staged", 43: "P flush start", 44: "P flush done",
         58: "I step begin", 59: "I operand ready", 60: "I wait rdy A", 61: "I wait rdy B", 62: "I rdy A", 63: "I rdy B",
         64: "I rows issued A", 65: "I rows issued B", 16: "fwd Dm read A", 17: "fwd Dm read B", 74: "I Dm barrier passed A", 75: "I Dm barrier passed B", 70: "I Dm wait A", 71: "I Dm wait B", 72: "I Dm free A", 73: "I Dm free B", 66: "I dW begin A", 67: "I dW begin B", 68: "I dW issued A", 69: "I dW issued B"}
path = sys.argv[1]
nrun = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mhz = float(sys.argv[3]) if len(sys.argv) > 3 else 1965.0
per = {}
for line in open(path):
    if line.startswith("TR "):
        _, w, e, t = line.split()
        per.setdefault(int(w), []).append((int(e), int(t)))
ev = []
for w, lst in per.items():
    n = len(lst) // nrun
    ev += [(t, w, e) for e, t in lst[-n:]]
ev.sort()
t0 = ev[0][0]
col = {0: 0, 16: 1, 20: 2}
tag = {0: "EW", 16: "PR", 20: "IS"}
for t, w, e in ev:
    print(f"{(t - t0) / mhz:8.3f} us  " + " " * (34 * col[w]) + f"[{tag[w]}] {NAMES.get(e, e)}")
