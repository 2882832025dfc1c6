#!/bin/bash
OUT=gpurun_out/r2l; mkdir -p $OUT
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_cuda_gpinn.py tests/test_cuda_widths.py -q -m gpu 2>&1 | tail -3 | tee $OUT/parity.txt
python -m pytest tests/test_cuda_solver.py -q -m gpu -k thousand -s 2>&1 | grep "fused-vs-fp64\|passed\|failed" | tee $OUT/thousand.txt
for st in trained trained_b; do PINN_TRAINED_DIR=$st PINN_LOAD_TRAINED=1 python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven Heat2D_Multiscale 2>&1 | tail -3; done | tee $OUT/acc.txt
bash scripts/whatif_times.sh r2l
