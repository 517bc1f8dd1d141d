"""CPU checks of the training-step oracle (oracle/training.py): TF1 Adam closed form, the TF BCE formula, the
variable split, gradient check of the restated graph by finite differences, and that the discriminator the
oracle trains is the one pinned by the golden vectors."""
import json
import math
import os

import numpy as np
import torch

from oracle import networks as on, training as ot

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_adam_tf1_closed_form():
    opt = ot.Adam(2e-4, 0.5)
    vals = {"w": np.array([1.0, -2.0, 0.5])}
    g = np.array([0.1, -0.2, 0.0])
    opt.step(vals, {"w": g})
    lr_t = 2e-4 * math.sqrt(1 - 0.999) / (1 - 0.5)
    m, v = 0.5 * g, 0.001 * g * g
    np.testing.assert_allclose(vals["w"], np.array([1.0, -2.0, 0.5]) - lr_t * m / (np.sqrt(v) + 1e-8), rtol=1e-12)
    opt.step(vals, {"w": g})
    assert opt.t == 2


def test_bce_matches_torch():
    x = torch.linspace(-30, 30, 41, dtype=torch.float64)
    for z in (0.0, 1.0):
        ref = torch.nn.functional.binary_cross_entropy_with_logits(x, torch.full_like(x, z), reduction="none")
        np.testing.assert_allclose(ot.bce_logits(x, z).numpy(), ref.numpy(), rtol=1e-9, atol=1e-12)


def test_variable_split_and_losses():
    cfg = on.make_cfg_4x(4, upRes=4, upsampling_mode=2, batch_norm=True)
    rng = np.random.default_rng(0)
    x, y = rng.random((3, 64), dtype=np.float32), rng.random((3, 256), dtype=np.float32)
    values = {}
    hp = dict(kk=5.0, kk2=1e-5, seed=2)
    ctx, L = ot.forward_losses(values, x, y, cfg, hp)
    train = [n for n, t in ctx.leaves.items() if t.requires_grad]
    assert all(ot.is_g_var(n) != ot.is_d_var(n) for n in train)  # the two substring filters partition the variables
    assert sum(ot.is_g_var(n) for n in train) == 12 * 2 + 9 * 2 and sum(ot.is_d_var(n) for n in train) == 5 * 2 + 3 * 2
    want = float(L["gen_loss"]) + 5.0 * float(L["gen_l1_loss"]) + 1e-5 * float(L["disc_loss_layer"])
    assert abs(float(L["gen_loss_complete"]) - want) < 1e-12
    # D step leaves the generator untouched and vice versa; BN moving statistics move in both
    before = {k: np.array(v, copy=True) for k, v in values.items()}
    ot.train_iteration(values, [(x, y)], [], cfg, hp, ot.Adam(2e-4, 0.5), ot.Adam(2e-4, 0.5))
    for n in before:
        changed = not np.array_equal(before[n], values[n])
        if n.endswith(("moving_mean", "moving_variance")):
            assert changed, n
        elif not ot.is_d_var(n):
            assert not changed, n
        elif n.endswith("weight"):  # (a bias in front of a batch norm has zero gradient and may stay put)
            assert changed, n


def test_discriminator_variables_match_reference_code():
    nets = np.load(os.path.join(GOLD, "nets.npz"))
    want = {k: tuple(v) for k, v in json.loads(str(nets["x4_mode2_disc_vars"]))}
    cfg = on.make_cfg_4x(16, upRes=4, upsampling_mode=2, batch_norm=True)
    values = {}
    rng = np.random.default_rng(1)
    ot.forward_losses(values, rng.random((1, 1024), dtype=np.float32), rng.random((1, 4096), dtype=np.float32), cfg,
                      dict(kk=1.0, kk2=1.0))
    have = {n: tuple(v.shape) for n, v in values.items() if n.startswith("discriminator/")}
    assert have == want
