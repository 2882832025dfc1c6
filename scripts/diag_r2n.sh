#!/bin/bash
OUT=gpurun_out/r2n; mkdir -p $OUT
timeout 900 python -m pytest tests/test_cuda_optim.py tests/test_cuda_solver.py tests/test_cuda_l2re.py tests/test_cuda_reference_model.py -q -m gpu 2>&1 | tail -4 | tee $OUT/tests.txt
for ts in 1 0; do echo "== PINN_TERM_STREAMS=$ts"; PINN_TERM_STREAMS=$ts python scripts/small_batch_latency.py 2>&1 | tail -8; done | tee $OUT/small_batch.txt
for ts in 1 0; do echo "== PINN_TERM_STREAMS=$ts c3 131072 rows (the 8-GPU share), eager"; PINN_C3_GRAPH=0 PINN_TERM_STREAMS=$ts python bench.py --workload c3 --points 131072 --steps 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['steps'], d['clocks']['sm_mhz'])"; done | tee $OUT/c3.txt
