export SKIP_REF=1 SKIP_CLASS=1 SKIP_NCU=1
bash scripts/gpu_round.sh r2final
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2final/smoke.txt 2>&1; tail -3 gpurun_out/r2final/smoke.txt
timeout 200 python scripts/sample_ncu.py > gpurun_out/r2final/sample_plain.txt 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -c 6 -o gpurun_out/r2final/ncu_sample -f python scripts/sample_ncu.py > gpurun_out/r2final/ncu_sample.log 2>&1
ncu -i gpurun_out/r2final/ncu_sample.ncu-rep --page raw --csv > gpurun_out/r2final/ncu_sample_raw.csv 2>/dev/null
ls -la gpurun_out/r2final
