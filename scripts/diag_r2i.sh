#!/bin/bash
OUT=gpurun_out/r2i; mkdir -p $OUT
for lib in pinnacle_b200/csrc/libpinn_b200.so variants/cheaplap/libpinn_b200.so; do
  echo "== $lib (state b)"; PINN_TRAINED_DIR=trained_b PINN_LOAD_TRAINED=1 PINN_LIB=$lib python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven Heat2D_Multiscale 2>&1 | tail -3
  echo "== $lib (state a)"; PINN_LOAD_TRAINED=1 PINN_LIB=$lib python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven Heat2D_Multiscale 2>&1 | tail -3
done | tee $OUT/acc.txt
python -m pytest tests/test_cuda_solver.py -q -m gpu -k thousand -s 2>&1 | grep "fused-vs-fp64\|passed\|failed" | tee $OUT/thousand.txt
bash scripts/whatif_times.sh r2i
