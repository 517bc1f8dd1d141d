#!/usr/bin/env python
"""Drop-in for GAN/multipassGAN-4x.py (output mode, `out 1`) of maxwerhahn/Multi-pass-GAN on the B200 path: same
`key value` flags, same input / output .uni files, one process per pass like GAN/example_run_output.py:6,8.
See multi-pass-gan_b200/cli_4x.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mpgan_b200  # noqa: E402,F401
from mpgan_b200 import cli_4x  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli_4x.main(sys.argv))
