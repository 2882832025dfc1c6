"""Hardware probe of the tcgen05 fp32 accumulate (run on a B200): how are the 16 products of one kind::f16 MMA and the
TMEM accumulator summed -- rounding mode, guard bits, alignment width.  Every operand value is exactly representable in
bf16 (and fp16 where that mode is probed), so whatever differs from the exact sum is the accumulate.

    python scripts/umma_round_probe.py [MODE]      MODE 10 = bf16 SS (default), 12 = bf16 TS, 20 = fp16 SS

Row m of A and column n of B are independent experiments: D[m, n] = sum_k A[m, k] B[k, n].
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinnacle_b200 import capi

MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lib = capi.load()
dev = torch.device("cuda:0")
N = 16
ULP2 = 2.0 ** -22          # spacing of fp32 in [2, 4)


def gemm(A, B):
    K = A.shape[1]
    At = torch.tensor(A, dtype=torch.float32, device=dev).contiguous()
    Bt = torch.tensor(B, dtype=torch.float32, device=dev).contiguous()
    D = torch.full((128, B.shape[1]), float("nan"), device=dev)
    rc = lib.pinn_debug_umma_gemm(MODE, 0, K, B.shape[1], ctypes.c_void_p(At.data_ptr()), ctypes.c_void_p(Bt.data_ptr()),
                                  ctypes.c_void_p(D.data_ptr()), None)
    torch.cuda.synchronize()
    assert rc == 0, rc
    return D.cpu().numpy().astype(np.float64)


fr = (np.arange(N) + 1) / 8.0          # 0.125 .. 2.0 ulp

print(f"== mode {MODE}")
# ---- E1: one MMA, big +-2 plus one small product f*ulp --------------------------------------------------------------
for K, tag, kpos in ((16, "E1 same MMA", 1), (32, "E3 accumulator from the previous k-step", 16)):
    A = np.zeros((128, K)); B = np.zeros((K, N))
    B[0, :] = 1.0; B[kpos, :] = fr * ULP2
    combos = [(+1, +1), (+1, -1), (-1, +1), (-1, -1)]
    for m, (sb, ss) in enumerate(combos):
        A[m, 0] = 2.0 * sb; A[m, kpos] = 1.0 * ss
    D = gemm(A, B)
    print(f"-- {tag}: D - (+-2) in ulp(2); columns f = 1/8 .. 16/8")
    for m, (sb, ss) in enumerate(combos):
        print(f"   big {sb:+d} small {ss:+d}: " + " ".join(f"{(D[m, n] - 2.0 * sb) / ULP2:+.2f}" for n in range(N)))

# ---- E2: guard bits: j small products that add up to f*ulp ----------------------------------------------------------
A = np.zeros((128, 16)); B = np.zeros((16, N))
B[0, :] = 1.0
js = [1, 2, 4, 8]
B[1:9, :] = 0.0
rows = []
for j in js:
    for sb in (+1, -1):
        m = len(rows)
        A[m, 0] = 2.0 * sb
        A[m, 1:1 + j] = sb * 1.0 / j          # j products of f*ulp/j each (1/j is a power of two)
        rows.append((j, sb))
B[1:9, :] = fr * ULP2
D = gemm(A, B)
print("-- E2 guard bits: big +-2 plus j equal small products summing to f*ulp (same sign as big); result in ulp")
for m, (j, sb) in enumerate(rows):
    print(f"   j={j} sign {sb:+d}: " + " ".join(f"{(D[m, n] - 2.0 * sb) / ULP2:+.2f}" for n in range(N)))

# ---- E4: alignment width: 2 - 2 + 2^-(20+n) -----------------------------------------------------------------------
tiny = 2.0 ** -(18.0 + 2 * np.arange(N))
for tag, K, layout in (("same MMA", 16, (0, 1, 2)), ("big pair in k-step 0, tiny in k-step 1", 32, (0, 1, 16)),
                       ("tiny in k-step 0, big pair in k-step 1", 32, (16, 17, 0)), ("+2 in k-step 0, -2 and tiny in k-step 1", 32, (0, 16, 17))):
    A = np.zeros((128, K)); B = np.zeros((K, N))
    A[0, layout[0]] = 2.0; A[0, layout[1]] = -2.0; A[0, layout[2]] = 1.0
    A[1, layout[0]] = -2.0; A[1, layout[1]] = 2.0; A[1, layout[2]] = 1.0
    A[2, layout[0]] = 2.0; A[2, layout[1]] = -2.0; A[2, layout[2]] = -1.0
    B[layout[0], :] = 1.0; B[layout[1], :] = 1.0; B[layout[2], :] = tiny
    D = gemm(A, B)
    print(f"-- E4 cancellation, {tag}: result / tiny for tiny = 2^-18, 2^-20, ... (1 = kept, 0 = lost)")
    for m in range(3):
        print("   " + " ".join(f"{D[m, n] / tiny[n]:+.3f}" for n in range(N)))

# ---- E5: exactness on a common grid (Ozaki-style operands): 8-bit integers, K = 112 --------------------------------
rng = np.random.default_rng(0)
K = 112
A = rng.integers(-127, 128, size=(128, K)).astype(np.float64)
B = rng.integers(-127, 128, size=(K, N)).astype(np.float64)
D = gemm(A, B)
ex = A @ B
print(f"-- E5 integer operands (|a|,|b| <= 127, K = 112): max |D - exact| = {np.abs(D - ex).max():.1f}, max |exact| = {np.abs(ex).max():.0f}")
A2 = A.copy(); A2[:, :16] *= 2.0 ** -12      # first k-step small terms (grid 2^-12), later k-steps integers
D = gemm(A2, B)
ex = A2 @ B
err = (D - ex)
ulp = 2.0 ** (np.floor(np.log2(np.abs(ex) + 1e-300)) - 23)
print(f"   small k-step first, then integers: error/ulp mean {np.mean(err / ulp):+.3f}  mean*sign(exact) {np.mean(err * np.sign(ex) / ulp):+.3f}  max |.| {np.abs(err / ulp).max():.2f}")
A3 = A.copy(); A3[:, -16:] *= 2.0 ** -12     # integers first, the last k-step adds small terms onto a big accumulator
D = gemm(A3, B)
ex = A3 @ B
err = (D - ex)
ulp = 2.0 ** (np.floor(np.log2(np.abs(ex) + 1e-300)) - 23)
print(f"   integers first, small k-step last: error/ulp mean {np.mean(err / ulp):+.3f}  mean*sign(exact) {np.mean(err * np.sign(ex) / ulp):+.3f}  max |.| {np.abs(err / ulp).max():.2f}")

# ---- E6: random bf16 operands, the statistics that matter: bias of the accumulate over 7 k-steps ---------------------
A = torch.randn(128, K).to(torch.bfloat16).double().numpy()
B = torch.randn(K, N).to(torch.bfloat16).double().numpy()
D = gemm(A, B)
ex = A @ B
ulp = 2.0 ** (np.floor(np.log2(np.abs(ex))) - 23)
err = (D - ex) / ulp
print(f"-- E6 random bf16 operands K = 112: error/ulp mean {err.mean():+.3f}  mean*sign(exact) {(err * np.sign(ex)).mean():+.3f}  rms {np.sqrt((err ** 2).mean()):.3f}  max {np.abs(err).max():.2f}")
for kk in (16, 32, 64):
    D = gemm(A[:, :kk], B[:kk])
    ex = A[:, :kk] @ B[:kk]
    ulp = 2.0 ** (np.floor(np.log2(np.abs(ex))) - 23)
    err = (D - ex) / ulp
    print(f"   K = {kk}: mean {err.mean():+.3f}  mean*sign {(err * np.sign(ex)).mean():+.3f}  rms {np.sqrt((err ** 2).mean()):.3f}")
