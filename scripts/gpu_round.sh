#!/bin/bash
# One gpurun call's worth of evidence: GPU tests, the default bench line, the reference arm, per-class kernel times,
# accuracy at trained weights, then (only after those exited without ncu) the ncu launch list and one full capture.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu ${PYTEST_X--x} -q --durations=15 > $OUT/pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
  # the numbers DESIGN.md section 5 quotes: L2RE of the replays, gradient error at the test's own trained states
  timeout 600 python -m pytest tests/test_cuda_l2re.py -m gpu -q -s 2>&1 | grep "L2RE" > $OUT/l2re.txt
  timeout 600 python -m pytest tests/test_cuda_solver.py -m gpu -q -s -k thousand 2>&1 | grep "fused-vs-fp64" > $OUT/trained_states.txt
fi
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; cat $OUT/bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err
  echo "ref rc=$?"; cat $OUT/bench_reference.json
fi
if [ -z "$SKIP_CLASS" ]; then
  timeout 600 python scripts/kernel_time.py --impls auto --modes fwd_bwd,fwd --problems Poisson2D_Classic,Heat2D_Multiscale,NS2D_LidDriven,Burgers1D,KuramotoSivashinskyEquation,GrayScottEquation,Poisson3D_ComplexGeometry,Heat2D_LongTime,Wave2D_Heterogeneous,PoissonND,HeatND > $OUT/per_class.txt 2>&1
  cat $OUT/per_class.txt
  PINN_LOAD_TRAINED=1 timeout 600 python scripts/accuracy_after_training.py Poisson2D_Classic Burgers1D Heat2D_Multiscale NS2D_LidDriven KuramotoSivashinskyEquation > $OUT/accuracy_trained.txt 2>&1
  cat $OUT/accuracy_trained.txt
fi
if [ -z "$SKIP_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $OUT/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pinn_tc2_kernel -s 2 -c 2 -o $OUT/ncu_full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $OUT/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ncu -i $OUT/ncu_full.ncu-rep --page raw --csv > $OUT/ncu_full_raw.csv 2>/dev/null
  ls -la $OUT
fi
