#!/bin/bash
# Device time of the two bench kernels for every what-if build under variants/*/libpinn_b200.so (scripts/kernel_time.py --lib).
OUT=gpurun_out/${1:-whatif}
mkdir -p $OUT
echo "== product build" | tee $OUT/times.txt
python scripts/kernel_time.py --impls auto --problems Poisson2D_Classic,Heat2D_Multiscale --iters 5 2>&1 | tee -a $OUT/times.txt
for lib in variants/*/libpinn_b200.so; do
  echo "== $lib" | tee -a $OUT/times.txt
  timeout 300 python scripts/kernel_time.py --lib $lib --impls auto --problems Poisson2D_Classic,Heat2D_Multiscale --iters 5 2>&1 | tail -4 | tee -a $OUT/times.txt
done
