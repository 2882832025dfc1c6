set -x
timeout -s KILL 400 python -m pytest tests/test_cuda_solver.py -x -q 2>&1 | tail -25
