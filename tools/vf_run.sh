timeout 900 python -m pytest tests/test_training8x_gpu.py -x -q 2>&1 | grep -E "Error|assert |passed|failed" | head -12
