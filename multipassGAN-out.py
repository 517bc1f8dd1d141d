#!/usr/bin/env python
"""Drop-in for GAN/multipassGAN-out.py of maxwerhahn/Multi-pass-GAN on the B200 path: same `key value` flags, same
input / output .uni files.  See multi-pass-gan_b200/cli.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mpgan_b200  # noqa: E402,F401
from mpgan_b200 import cli  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli.main(sys.argv))
