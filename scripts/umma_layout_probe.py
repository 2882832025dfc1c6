"""Layout discovery: A selects row k (A[m][k] = delta), B buffer holds its own word offsets -> D[k][n] = offset read for (k, n)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinnacle_b200 import capi
variant, K, N = (int(a) for a in sys.argv[1:4])
lib = capi.load()
dev = torch.device("cuda:0")
A = torch.zeros(128, K, device=dev)
for k in range(K):
    A[k, k] = 1.0
B = torch.zeros(K, N, device=dev)
D = torch.full((128, N), float("nan"), device=dev)
rc = lib.pinn_debug_umma_gemm(4, variant, K, N, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(D.data_ptr()), None)
torch.cuda.synchronize()
print(f"variant={variant} K={K} N={N} rc={rc}")
for k in range(min(K, 16)):
    print(f"k={k:2d}:", " ".join(f"{int(v):5d}" for v in D[k].tolist()))
