set -x
timeout -s KILL 120 python scripts/kernel_time.py --problems Poisson2D_Classic --impls tc2,tc --modes fwd_bwd --check --points 100000 || echo "KERNEL HUNG OR FAILED"
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 200 python scripts/kernel_time.py --problems Poisson2D_Classic,Heat2D_Multiscale,NS2D_LidDriven,KuramotoSivashinskyEquation,Burgers1D,Poisson3D_ComplexGeometry,Heat2D_LongTime,NS2D_LongTime,Burgers2D --impls auto --modes fwd_bwd
