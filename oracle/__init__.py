"""ORACLE - test infrastructure only.

CPU restatement (torch-CPU fp32/fp64 + numpy/scipy) of the reference algorithm for the generator hot
path of maxwerhahn/Multi-pass-GAN.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference legs may import this package; the product path (multi-pass-gan_b200/) never does and fails
loudly when its CUDA library is missing.

PARITY UNPINNED: the reference has no tests, golden vectors or fixtures for this path
and TensorFlow 1.x / Keras cannot be installed in this image, so the oracle restates the published op
semantics and follows the reference call sites cited in every docstring.  Its own pins
are the known-answer tests in tests/test_oracle_*.py and the committed vectors in tests/golden/.
"""
