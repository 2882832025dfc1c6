#!/bin/bash
# diagnostics: which side goes NaN in the KS Adam->L-BFGS case; trained-weights error with the two tanh variants
OUT=gpurun_out/r2d; mkdir -p $OUT
for i in 1 2; do python tests/ref_gpu_driver.py train KuramotoSivashinskyEquation lbfgs 50 2>&1 | grep RESULT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[7:]); import numpy as np
    hr=np.array(d['hist_reference']); hf=np.array(d['hist_fused'])
    print('reference finite', np.isfinite(hr).all(), 'fused finite', np.isfinite(hf).all()); print('ref', hr[:, :2].tolist()); print('fused', hf[:, :2].tolist()); print(d['hist_rel'], d['weights_rel'])
" ; done > $OUT/ks_lbfgs.txt 2>&1
cat $OUT/ks_lbfgs.txt
python -m pytest tests/test_cuda_solver.py -q -m gpu -k thousand -s 2>&1 | grep "fused-vs-fp64\|passed\|failed" | tee $OUT/thousand_product.txt
for lib in variants/*/libpinn_b200.so; do echo "== $lib"; PINN_B200_LIB=$PWD/$lib python -m pytest tests/test_cuda_solver.py -q -m gpu -k thousand -s 2>&1 | grep "fused-vs-fp64\|passed\|failed"; done | tee $OUT/thousand_variants.txt
bash scripts/whatif_times.sh r2d
for lib in pinnacle_b200/csrc/libpinn_b200.so variants/evictlast/libpinn_b200.so; do
  echo "== dram $lib"; timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pinn_tc2 -c 2 python scripts/kernel_time.py --lib $lib --impls auto --problems Poisson2D_Classic,Heat2D_Multiscale --iters 1 --modes fwd_bwd 2>&1 | grep -E "dram__|gpu__time|pinn_tc2_kernel" ; done | tee $OUT/dram.txt
