"""Host-side schedules of the 8x progressive-growing trainer (GAN/multipassGAN-8x.py), no device code:

* `GrowthSchedule` -- which data resolution (`currentUpres`), which per-stage optimizer (`index`) and which blending
  `percentage` every iteration of the training loop uses, and the learning-rate step counter `lrgs`
  (:211-218 initial resolution, :1884-1896 initial blend counters incl. resumed runs (`startingIter`), :1905-1916 growing
  events, :1965-1982 blend value / optimizer index / decay counter).  Pinned by golden vectors traced by executing the
  reference's own statements (tests/golden/make_golden.py schedule -> schedule8x.npz).
* `polynomial_decay` / `learning_rates` -- tf.train.polynomial_decay(lr, lr_global_step, decayIter, lr * 0.05, power=1.1)
  as the reference configures it (:995-1017).
* `zero_density_batch` -- the 1-in-20 "empty input" batches of getinput (:1527-1533), same numpy random calls.

Reference behaviour kept as is (it looks unintended, the trainer reproduces it instead of repairing it):
the blend counter starts one stage ahead (`interpol_c += stageIter`, :1893), so `percentage` already runs from 1 to 2
while the data of the first stage (currentUpres 2) is trained and ends at 3 one stage before the data reaches 8x.
For the refinement networks (upsampling_mode 1 / 3) there are no growing events, so the counter stops after the first blend
phase and `percentage` stays at 2: under the reference's own schedule their 8x stage is never blended in.
"""
import math
from collections import namedtuple

import numpy as np

Step = namedtuple("Step", "it currentUpres index percentage lrgs grew")


class GrowthSchedule:
    """Iterable over the training iterations `startingIter .. 6 * stageIter + decayIter - 1` (:166, :1898)."""

    def __init__(self, stageIter=25000, decayIter=25000, upRes=8, upsampling_mode=2, startingIter=0, decayLR=True):
        self.stageIter, self.decayIter, self.upRes = int(stageIter), int(decayIter), int(upRes)
        self.mode, self.startingIter, self.decayLR = int(upsampling_mode), int(startingIter), bool(decayLR)
        if self.stageIter < 1:
            raise ValueError("stageIter must be >= 1")
        self.trainingIterations = self.stageIter * 6 + self.decayIter

    def initial_upres(self):
        """Resolution of the training data at `startingIter`: grows 2 -> 4 -> 8 for the first network (upsampling_mode 2),
        fixed 8 for the refinement networks (:211-218)."""
        if self.mode == 2:
            return min(2 ** (self.startingIter // (self.stageIter * 2) + 1), 8)
        return 8

    def __len__(self):
        return max(0, self.trainingIterations - self.startingIter)

    def __iter__(self):
        s, t0 = self.stageIter, self.startingIter
        pair = t0 // (2 * s)                 # finished (blend, stabilise) stage pairs
        blend_end = s * pair * 2 + s         # iteration at which the running blend phase stops counting
        count = pair * s                     # blend counter; percentage = count / stageIter
        blending = (t0 // s) % 2 == 0
        if blending:
            count += (t0 - count) % s + s
        upres, lrgs = self.initial_upres(), 0
        for it in range(t0, self.trainingIterations):
            if it == blend_end:
                blending = False
            grew = False
            if it - blend_end == s and upres < self.upRes:
                upres *= 2
                blend_end = it + s
                blending, grew = True, True
            if blending:
                count += 1
                pct = count / s
            else:
                pct = int(round(count / s))
            pct = min(max(pct, 1.0), 3.0)
            if it >= s * 6 and self.decayLR:
                lrgs += 1
            yield Step(it, upres, int(round(math.log(upres, 2)) - 1), pct, lrgs, grew)


def polynomial_decay(learning_rate, global_step, decay_steps, end_learning_rate, power):
    """tf.train.polynomial_decay (cycle=False): (lr - end) * (1 - min(step, decay_steps) / decay_steps) ** power + end."""
    step = min(float(global_step), float(decay_steps))
    return (learning_rate - end_learning_rate) * (1.0 - step / float(decay_steps)) ** power + end_learning_rate


def learning_rates(learning_rate, lrgs, decayIter, decayLR=True, n_stages=3):
    """Per-stage learning rates (generator list, discriminator list) of :995-1017. With decayLR every stage uses
    polynomial_decay(lr, lrgs, decayIter, 0.05 * lr, power=1.1). Without it the reference's generator list is [lr] * 3 and
    its discriminator list has ONE entry (lr / 4, appended at i == 1 only), so building `disc_optimizer[1]` fails there
    (IndexError, :1340): the same configuration is refused here."""
    if decayLR:
        lr = polynomial_decay(learning_rate, lrgs, decayIter, learning_rate * 0.05, 1.1)
        return [lr] * n_stages, [lr] * n_stages
    raise ValueError("decayLR 0: the reference builds a single discriminator learning rate for three optimizers "
                     "(GAN/multipassGAN-8x.py:1010-1017) and cannot run; use decayLR 1")


def zero_density_batch(batch_xs, batch_ys, add_adj_idcs, rng=np.random):
    """getinput :1527-1533: with probability 1/20 (`np.random.randint(0, 20) == 0`) the batch becomes an empty-density one:
    density (and the adjacent-slice densities, channels 4:6) zeroed, velocities scaled by 1 + 1.5 * np.random.rand(), targets
    zeroed.  Arrays are [..., C] with the channel last (numpy or torch); modified in place.  Returns whether it happened."""
    if min(rng.randint(0, 20), 1):
        return False
    batch_xs[..., 0:1] = 0
    if add_adj_idcs:
        batch_xs[..., 4:6] = 0
    batch_xs[..., 1:4] *= (1.0 + rng.rand() * 1.5)
    batch_ys[...] = 0
    return True
