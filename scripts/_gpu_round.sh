set -x
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python scripts/kernel_time.py --problems Wave2D_Heterogeneous,Poisson2D_Classic,Heat2D_Multiscale,NS2D_LidDriven,Heat2D_LongTime --impls auto --modes fwd_bwd
