timeout 900 python -m pytest tests/test_training_gpu.py -x -q 2>&1 | tail -3
