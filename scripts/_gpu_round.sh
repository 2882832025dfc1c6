set -x
timeout -s KILL 400 python -m pytest tests/test_cuda_optim.py -x -q 2>&1 | tail -12
