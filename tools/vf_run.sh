timeout 900 python -m pytest tests/test_training_gpu.py tests/test_plumbing_gpu.py tests/test_training8x_gpu.py -x -q 2>&1 | tail -3
python tools/train_launch_hist.py 2>&1 | grep -A8 "device time"
python tools/bench_train.py 2>&1 | tail -1 | cut -c1-330
