set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_final.json 2>gpurun_out/bench_ref_err.log
python bench.py > gpurun_out/r1_bench_final.json 2>gpurun_out/bench_err.log; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pinn_tc2_kernel -s 2 -c 2 -o gpurun_out/r1_final_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
python scripts/kernel_time.py --problems Burgers1D,Burgers2D,Poisson2D_Classic,PoissonBoltzmann2D,Poisson3D_ComplexGeometry,PoissonND,Heat2D_Multiscale,Heat2D_ComplexGeometry,Heat2D_LongTime,HeatND,NS2D_LidDriven,NS2D_BackStep,NS2D_LongTime,Wave1D,Wave2D_LongTime,KuramotoSivashinskyEquation,GrayScottEquation,PoissonInv,HeatInv --impls auto --modes fwd_bwd,fwd > gpurun_out/per_class_final.txt 2>&1
