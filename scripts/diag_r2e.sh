#!/bin/bash
OUT=gpurun_out/r2e; mkdir -p $OUT
rm -f gpurun_out/trained_*.pt
python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven 2>&1 | tail -2 | tee $OUT/acc_state.txt
for lib in variants/*/libpinn_b200.so; do echo "== $lib"; PINN_LOAD_TRAINED=1 PINN_LIB=$lib python scripts/accuracy_after_training.py Poisson2D_Classic NS2D_LidDriven 2>&1 | tail -2; done | tee -a $OUT/acc_state.txt
bash scripts/whatif_times.sh r2e
python tests/ref_gpu_driver.py train KuramotoSivashinskyEquation lbfgs 35 2>&1 | grep RESULT | cut -c1-1500 | tee $OUT/ks35.txt
