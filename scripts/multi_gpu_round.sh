#!/bin/bash
# N-GPU evidence: NCCL tests + the bench line at N GPUs (torchrun, as the driver launches it).   gpurun --gpus N -- 'bash scripts/multi_gpu_round.sh N tag'
N=${1:-2}; TAG=${2:-r2_multi}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_cuda_multigpu.py -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_multigpu.txt
fi
timeout ${BENCH_TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
echo "bench rc=$?"; tail -c 600 $OUT/bench_${N}gpu.err; python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_${N}gpu.json") if l.startswith("{")][-1])
print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
for k,v in d["config"].get("extra_workloads",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("steps"), v.get("clocks"), v.get("config",{}).get("cuda_graph"), v.get("error"))
PY
