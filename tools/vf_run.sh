for g in 2 3 4; do echo "== groups $g"; MPG_VRING_GROUPS=$g python tools/prof_conv.py n2_48to48,n2_48_96to48,n1_64to64k3 4; done
MPG_VRING_GROUPS=4 timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -k "vr_ or vring" 2>&1 | tail -2
