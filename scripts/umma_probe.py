"""Hardware probe of the tcgen05 descriptor conventions (run on a B200): python scripts/umma_probe.py MODE VARIANT K N"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinnacle_b200 import capi

mode, variant, K, N = (int(a) for a in sys.argv[1:5])
lib = capi.load()
torch.manual_seed(0)
dev = torch.device("cuda:0")


def tf32(x):  # round to 10-bit mantissa (bf16 for the 16-bit modes)
    if mode >= 10:
        return x.to(torch.bfloat16).float()
    i = x.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


A = tf32(torch.randn(128, K, device=dev))
B = tf32(torch.randn(K, N, device=dev))
D = torch.full((128, N), float("nan"), device=dev)
rc = lib.pinn_debug_umma_gemm(mode, variant, K, N, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                              ctypes.c_void_p(D.data_ptr()), None)
torch.cuda.synchronize()
ref = (A.double() @ B.double())
err = float((D.double() - ref).abs().max() / ref.abs().max())
print(f"mode={mode} variant={variant} K={K} N={N} rc={rc} rel-max-err={err:.3e} nan={int(torch.isnan(D).sum())}")
