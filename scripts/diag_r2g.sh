#!/bin/bash
OUT=gpurun_out/r2g; mkdir -p $OUT
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_cuda_gpinn.py tests/test_cuda_widths.py tests/test_cuda_evaluate.py -q -m gpu 2>&1 | tail -8 | tee $OUT/parity.txt
python -m pytest tests/test_cuda_solver.py -q -m gpu -k thousand -s 2>&1 | grep "fused-vs-fp64\|passed\|failed" | tee $OUT/thousand.txt
PINN_LOAD_TRAINED=1 python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven Heat2D_Multiscale Burgers1D KuramotoSivashinskyEquation 2>&1 | tail -5 | tee $OUT/acc_state.txt
bash scripts/whatif_times.sh r2g
