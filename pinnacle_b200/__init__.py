"""pinnacle_b200 -- fused sm_100a PDE-residual + parameter-gradient step for PINNacle.

Host side is Python/PyTorch; the compute path is ``csrc/libpinn_b200.so`` (hand-written CUDA
behind a C ABI, ``include/pinn_b200.h``).  There is no CPU fallback: importing the op without
the built library raises.
"""
from . import kinds, problems  # noqa: F401

__version__ = "0.1.0"
