"""Schedules of the 8x trainer: GrowthSchedule / zero_density_batch against traces produced by executing the
reference's own statements (GAN/multipassGAN-8x.py:211-218, 1884-1896, 1905-1982, 1527-1533; make_golden.py schedule)."""
import json
import os

import numpy as np
import pytest

from mpgan_b200 import schedule8x as S

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "schedule8x.npz"))
CASES = sorted(k[:-4] for k in GOLD.files if k.endswith("_cfg"))


@pytest.mark.parametrize("tag", CASES)
def test_growth_schedule_reproduces_the_reference_loop(tag):
    c = json.loads(str(GOLD[tag + "_cfg"]))
    sch = S.GrowthSchedule(c["stageIter"], c["decayIter"], c["upRes"], c["upsampling_mode"], c["startingIter"], c["decayLR"])
    want = GOLD[tag]
    got = list(sch)
    assert len(got) == len(want) == len(sch)
    grown = 0
    for st, w in zip(got, want):
        grown += int(st.grew)
        assert (st.it, st.currentUpres, st.index, st.lrgs) == (int(w[0]), int(w[1]), int(w[2]), int(w[4])), (tag, st, w)
        assert st.percentage == w[3], (tag, st, w)          # same float arithmetic: exact
        assert 2 * grown == int(w[5])                        # one copyAdamVariables + one saveModel per growing event


def test_cases_cover_growing_resume_and_refinement_modes():
    assert any(GOLD[t][:, 5].max() == 4 for t in CASES) and any("resume" in t for t in CASES) and any(t.startswith("m1") for t in CASES)


def test_polynomial_decay_known_answers():
    lr = 1e-4
    assert S.polynomial_decay(lr, 0, 100, lr * 0.05, 1.1) == pytest.approx(lr)
    assert S.polynomial_decay(lr, 100, 100, lr * 0.05, 1.1) == pytest.approx(lr * 0.05)
    assert S.polynomial_decay(lr, 250, 100, lr * 0.05, 1.1) == pytest.approx(lr * 0.05)   # clamped at decay_steps
    mid = S.polynomial_decay(lr, 50, 100, lr * 0.05, 1.1)
    assert mid == pytest.approx(0.95 * lr * 0.5 ** 1.1 + 0.05 * lr)
    g, d = S.learning_rates(lr, 50, 100)
    assert g == [mid] * 3 and d == [mid] * 3
    with pytest.raises(ValueError):
        S.learning_rates(lr, 0, 100, decayLR=False)


def test_zero_density_batches_match_the_reference_statement():
    hits = set(GOLD["zero_hits"].tolist())
    assert 1 <= len(hits) <= 10
    for seed in range(60):
        xs, ys = GOLD["zero_xs0"].copy(), GOLD["zero_ys0"].copy()
        np.random.seed(seed)
        hit = S.zero_density_batch(xs, ys, True)
        assert hit == (seed in hits)
        if seed == int(GOLD["zero_seed"]):
            assert np.array_equal(xs, GOLD["zero_xs"]) and np.array_equal(ys, GOLD["zero_ys"])
        if not hit:
            assert np.array_equal(xs, GOLD["zero_xs0"]) and np.array_equal(ys, GOLD["zero_ys0"])
