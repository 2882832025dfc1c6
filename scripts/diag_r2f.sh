#!/bin/bash
OUT=gpurun_out/r2f; mkdir -p $OUT
rm -f gpurun_out/trained_*.pt
python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven Heat2D_Multiscale Burgers1D KuramotoSivashinskyEquation 2>&1 | tail -5 | tee $OUT/acc_state.txt
python -m pytest tests/test_cuda_solver.py -q -m gpu -k thousand -s 2>&1 | grep "fused-vs-fp64\|passed\|failed" | tee $OUT/thousand.txt
bash scripts/whatif_times.sh r2f
