set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:pinn_tc2 -s 2 -c 2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -A8 "pinn_tc2_kernel" | grep "pinn_tc2\|dram\|duration"
python bench.py --workload c5 --points 16777216 --scaling strong --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_c5_16M_1gpu.json 2>gpurun_out/bench_err.log; cut -c1-400 gpurun_out/r1_bench_c5_16M_1gpu.json
