#!/usr/bin/env python
"""Drop-in for GAN/multipassGAN-8x.py in TRAINING mode (`out 0`, first network) of maxwerhahn/Multi-pass-GAN on the B200 path:
same `key value` flags, the same .uni simulation files in, `basePath/test_%04d/model[_ema]_%04d.ckpt` out (restored by
multipassGAN-out.py). See multi-pass-gan_b200/cli_8x.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mpgan_b200  # noqa: E402,F401
from mpgan_b200 import cli_8x  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli_8x.main(sys.argv))
