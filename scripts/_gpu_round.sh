set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/kernel_time.py --problems Poisson2D_Classic,Heat2D_Multiscale,NS2D_LidDriven,KuramotoSivashinskyEquation,Burgers1D,Poisson3D_ComplexGeometry,Heat2D_LongTime,PoissonND,Wave2D_Heterogeneous --impls auto --modes fwd_bwd,fwd
