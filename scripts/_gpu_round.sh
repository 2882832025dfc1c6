set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_lap.json 2>gpurun_out/bench_ref_err.log
python bench.py > gpurun_out/r1_bench_lap.json 2>gpurun_out/bench_err.log; echo bench rc=$?
cut -c1-600 gpurun_out/r1_bench_lap.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_lap.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pinn_tc2_kernel -s 2 -c 2 -o gpurun_out/r1_lap_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_lap.log 2>&1
python scripts/kernel_time.py --problems Poisson2D_Classic,Heat2D_Multiscale,Poisson3D_ComplexGeometry,Heat2D_LongTime,Wave2D_Heterogeneous,KuramotoSivashinskyEquation,Burgers1D,NS2D_LidDriven --impls auto --modes fwd_bwd,fwd
