"""Convert between TF1 checkpoints (checkpoint V2: <prefix>.index + <prefix>.data-00000-of-00001) and .npz archives,
without TensorFlow (multi-pass-gan_b200/tfckpt.py).

  python tools/ckpt_convert.py list   model_0009.ckpt
  python tools/ckpt_convert.py to-npz model_0009.ckpt weights.npz [--scope gen_1]   # keys get the scope prefix the
                                                                                    # apply CLI's `weightsNpz` expects
  python tools/ckpt_convert.py to-ckpt weights.npz model_0009.ckpt [--strip gen_1]  # inverse (strip the scope)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

import mpgan_b200  # noqa: F401
from mpgan_b200 import tfckpt


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("list")
    p.add_argument("prefix")
    p = sub.add_parser("to-npz")
    p.add_argument("prefix")
    p.add_argument("npz")
    p.add_argument("--scope", default="", help="prefix every key with <scope>/ (e.g. gen_1)")
    p = sub.add_parser("to-ckpt")
    p.add_argument("npz")
    p.add_argument("prefix")
    p.add_argument("--strip", default="", help="only keys under <strip>/ are written, without that prefix")
    a = ap.parse_args(argv)
    if a.cmd == "list":
        for name, e in sorted(tfckpt.list_checkpoint(a.prefix).items()):
            if name:
                print("%-70s dtype %-3d shape %s" % (name, e["dtype"], tuple(e["shape"])))
        return 0
    if a.cmd == "to-npz":
        pre = a.scope + "/" if a.scope else ""
        np.savez(a.npz, **{pre + k: v for k, v in tfckpt.read_checkpoint(a.prefix, verify_data=True).items()})
        return 0
    arch = np.load(a.npz)
    pre = a.strip + "/" if a.strip else ""
    tensors = {k[len(pre):]: arch[k] for k in arch.files if k.startswith(pre)}
    if not tensors:
        raise SystemExit("no keys under '%s' in %s" % (pre, a.npz))
    tfckpt.write_checkpoint(a.prefix, tensors)
    return 0


if __name__ == "__main__":
    sys.exit(main())
