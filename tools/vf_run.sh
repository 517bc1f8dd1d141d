timeout 600 python -m pytest tests/test_training8x_gpu.py -x -q 2>&1 | tail -25
