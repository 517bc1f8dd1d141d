timeout 900 python -m pytest tests/test_training_gpu.py tests/test_training8x_gpu.py -x -q 2>&1 | tail -3
python tools/bench_train.py 2>/dev/null | tail -1 | cut -c1-260
TIMED=1 python tools/train_launch_hist.py 2>&1 | grep -A10 "device time"
