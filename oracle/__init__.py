"""ORACLE - test infrastructure only.

CPU restatement (torch-CPU fp32/fp64 + numpy/scipy) of the reference algorithm for the generator hot
path of maxwerhahn/Multi-pass-GAN.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference legs may import this package; the product path (multi-pass-gan_b200/) never does and fails
loudly when its CUDA library is missing.

PARITY PINNING: the reference has no tests, golden vectors or fixtures for this path
and TensorFlow 1.x / Keras cannot be installed in this image.  What IS pinned, by executing the
reference's own Python here (tests/golden/make_golden.py -> tests/golden/*.npz, checked by
tests/test_golden.py):
  * the volume pipeline (generate3DUniForNewNetwork of multipassGAN-out.py / multipassGAN-4x.py, run
    with stand-in row functions for `sess.run`): oracle/pipeline.py reproduces it bit-exactly;
  * layer wiring, variable names/shapes, wscale constants, BN/bias/activation order and the cursor
    quirks (the real tools_wscale/GAN.py class + growing_gen / gen_resnet / disc_binclass executed on a
    numpy TF1 shim): the fp64 oracle reproduces those outputs to 1e-9;
  * the 8x progressive-growing trainer's nets (growing_disc / growBlockDisc / lerp and growing_gen in TRAINING mode of
    multipassGAN-8x.py, same shim; tests/golden/growdisc.npz, tests/test_oracle_training8x.py): oracle/training8x.py
    reproduces logits, feature layers, blended outputs and variable names; gradients (incl. the WGAN-GP double backward)
    are torch autograd in fp64 on that restatement.
What stays UNPINNED (restated from the published TF semantics): the arithmetic of
tf.nn.conv2d SAME, contrib batch_norm and resize_images (nearest / TF1 legacy bicubic) themselves -
the shim and oracle/tf_ops.py are two independent restatements that agree, not TensorFlow output.
"""
