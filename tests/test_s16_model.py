"""The number format of the next fill kernel, checked before any device code exists: the linear
GuidedAlign sweep with 16-bit slots relative to an offset re-based every 64 anti-diagonals gives every in-band cell the
arrow and the score of the 32-bit sweep, whose end score is pinned against the oracle (orc_align / the reference)."""
import numpy as np
import pytest

from blasr_b200 import SMRTDistanceMatrix
from tests import cases, oracle as O

WHICH = "ref" if O.have_ref() else "orc"


@pytest.mark.parametrize("at", [1, 0])          # Global, Local
def test_s16_slots_reproduce_the_32_bit_sweep(at):
    rng = np.random.default_rng(31 + at)
    worst = {"min_rel": 0, "max_rel": 0, "max_big": 0}
    cells = 0
    for rep, (lo, hi, n) in enumerate(((100, 1500, 24), (3000, 12000, 6), (15000, 20000, 2))):
        b = cases.guided_batch(seed=400 + rep, n=n, lo=lo, hi=hi, n_rate=0.003)
        for i in range(b.n):
            q, t, g, _ = cases.job_arrays(b, i)
            band = int(rng.choice([8, 16, 32, 64]))
            ins, dele = (5, 5) if i % 2 == 0 else (int(rng.integers(1, 10)), int(rng.integers(1, 10)))
            fn = O.score_fn(SMRTDistanceMatrix, ins, dele)
            j, keep = O.make_job(0, at, band, q, t, g, None, 0, 0, 0, 0)
            m = O.guided_s16_model(fn, j)
            assert m is not None
            want = O.align(WHICH, fn, j)
            assert want["status"] == 0
            assert m["end32"] == want["score"], (rep, i, m["end32"], want["score"])       # the 32-bit sweep is GuidedAlign
            assert m["end16"] == want["score"]
            assert m["arrow_mismatches"] == 0 and m["score_mismatches"] == 0, (rep, i, band, m)
            assert m["cells"] == want["nCells"] - max(0, 0) or m["cells"] <= want["nCells"]   # cells past tEnd are never filled
            cells += m["cells"]
            worst["min_rel"] = min(worst["min_rel"], m["min_rel"]); worst["max_rel"] = max(worst["max_rel"], m["max_rel"])
            worst["max_big"] = max(worst["max_big"], m["max_big"])
    assert cells > 3_000_000
    # the format's head-room on these workloads: legit values far below the threshold, unreachable ones inside int16
    assert worst["max_rel"] < 20000 and worst["min_rel"] > -12000 and worst["max_big"] < 32767, worst


def test_s16_model_on_adversarial_guides():
    """Anchor-only guides with long gaps (wide windows, large in-window score spread)."""
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5)
    b = cases.guided_batch(seed=77, n=24, lo=200, hi=3000, adversarial=0.5, run=12)
    checked = 0
    for i in range(b.n):
        q, t, g, _ = cases.job_arrays(b, i)
        j, keep = O.make_job(0, 1, 16, q, t, g, None, 0, 0, 0, 0)
        want = O.align(WHICH, fn, j)
        m = O.guided_s16_model(fn, j)
        if want["status"] != 0 or m is None:
            continue
        assert m["end32"] == want["score"]
        # wide post-gap rows can push the in-window spread past the 16-bit budget: the model must SAY so (score_mismatches > 0)
        # rather than silently disagree, and then the kernel has to route the job to the 32-bit path
        if m["score_mismatches"] == 0:
            assert m["end16"] == want["score"] and m["arrow_mismatches"] == 0
        checked += 1
    assert checked >= 12
