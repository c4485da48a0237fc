"""The oracle as a model: golden fixtures, replay chain, layout equivalence, analytic results.

Known-answer tests the reference does not hold (SURVEY.md 8c "known-answer tests the build must
create itself"): invariants of the process plus statistical equivalence of the reference layout
(per-cell vector, ChaCha8/ziggurat/BTPE) and the GPU layout (histogram, Philox/popcount).
"""
import json
import os
import sys

import numpy as np
import pytest
from scipy import stats as sps

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "golden_v2.json")))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_matches_golden_fixtures(name):
    kw = make_golden.CASES[name]
    for i, want in enumerate(GOLD["native"][name]):
        got = make_golden.record(ob.run(ob.make_opts(state=ob.STATE_HIST, rng=ob.RNG_PHILOX, run_idx=260 + i, **kw),
                                        hist_cap=512))
        assert got == want
    for i, want in enumerate(GOLD["reference_layout"][name]):
        got = make_golden.record(ob.run(ob.make_opts(state=ob.STATE_VECTOR, rng=ob.RNG_RAND, run_idx=260 + i, **kw),
                                        hist_cap=512))
        assert got == want


@pytest.mark.parametrize("rng", [ob.RNG_RAND, ob.RNG_PHILOX])
def test_replay_reproduces_every_event(rng):
    """vector oracle -> decision stream -> histogram oracle: identical state after every event."""
    kw = dict(b0=1.0, b1=1.2, d0=0.2, d1=0.25, max_cells=3000, initial={3: 5, 0: 2})
    src = ob.run(ob.make_opts(state=ob.STATE_VECTOR, rng=rng, **kw), hist_cap=512, trace_cap=100000, traj_cap=100000)
    assert src.trace_len == src.n_events
    rep = ob.run(ob.make_opts(state=ob.STATE_HIST, rng=ob.RNG_REPLAY, replay=src.trace, **kw), hist_cap=512,
                 traj_cap=100000)
    assert rep.n_events == src.n_events and rep.stop_reason == src.stop_reason
    np.testing.assert_array_equal(rep.traj, src.traj)  # nminus, nplus, clock bits, histogram digest per event
    np.testing.assert_array_equal(rep.hist, src.hist)
    assert (rep.hash, rep.chain) == (src.hash, src.chain)
    # a truncated stream stops with REPLAY_END exactly there
    cut = ob.run(ob.make_opts(state=ob.STATE_HIST, rng=ob.RNG_REPLAY, replay=src.trace[:100], **kw), hist_cap=512)
    assert cut.stop_reason == ob.STOP_REPLAY_END and cut.n_events == 100


def test_pure_birth_event_count_and_conservation():
    r = ob.run(ob.make_opts(max_cells=5000, state=ob.STATE_VECTOR, rng=ob.RNG_RAND), hist_cap=512)
    assert r.stop_reason == ob.STOP_MAX_CELLS and r.n_events == 4999 and r.nminus + r.nplus == 5000
    assert r.hist.sum() == 5000 and r.hist[0] == r.nminus


@pytest.mark.parametrize("state", [ob.STATE_VECTOR, ob.STATE_HIST])
def test_deterministic_rule_keeps_copy_number(state):
    r = ob.run(ob.make_opts(b1=1.0, max_cells=2000, segregation=ob.SEG_DETERMINISTIC, initial={7: 1}, state=state,
                            rng=ob.RNG_PHILOX), hist_cap=64)
    assert r.nminus == 0 and r.hist[7] == 2000 and r.hist.sum() == 2000 and r.kmax == 7


@pytest.mark.parametrize("rng", [ob.RNG_RAND, ob.RNG_PHILOX])
def test_no_uneven_never_creates_nminus(rng):
    r = ob.run(ob.make_opts(b0=0.0, b1=1.0, max_cells=3000, segregation=ob.SEG_BINOMIAL_NO_UNEVEN,
                            state=ob.STATE_VECTOR, rng=rng), hist_cap=512)
    assert r.nminus == 0 and r.nplus == 3000


def test_no_nminus_rule_event_count():
    """BinomialNoNminus: an uneven split adds no cell, so events >= cells - 1."""
    r = ob.run(ob.make_opts(b0=1.0, b1=1.0, max_cells=2000, segregation=ob.SEG_BINOMIAL_NO_NMINUS), hist_cap=512)
    assert r.nminus + r.nplus == 2000 and r.n_events > 1999


def test_stop_rules():
    assert ob.run(ob.make_opts(max_cells=10 ** 6, max_time=3.0)).stop_reason == ob.STOP_MAX_TIME
    assert ob.run(ob.make_opts(max_cells=10 ** 6, max_iter=50)).n_events == 49  # iter >= max_iter - 1
    r = ob.run(ob.make_opts(b0=0.0, b1=0.0, max_cells=100, initial={0: 5}))
    assert r.stop_reason == ob.STOP_ABSORBING and r.n_events == 0
    r = ob.run(ob.make_opts(b0=0.0, b1=0.0, d0=1.0, d1=1.0, max_cells=100, initial={2: 3, 0: 4}))
    assert r.stop_reason == ob.STOP_NO_INDIVIDUALS and r.n_events == 7
    r = ob.run(ob.make_opts(max_cells=100, initial={40000: 1}, segregation=ob.SEG_DETERMINISTIC), hist_cap=64)
    assert r.stop_reason == ob.STOP_COPY_OVERFLOW
    # sosa-sum counting mode: the birth-death population array is [n-,n+,n-,n+]
    a = ob.run(ob.make_opts(b1=1.2, d0=0.1, d1=0.1, max_cells=1000, bd_count_mode=1, initial={1: 30}))
    assert a.stop_reason == ob.STOP_MAX_CELLS and a.nminus + a.nplus == 500


def test_snapshot_front_pop_quirk():
    """process.rs:122-129: ANY remaining size matching pops the FRONT entry; an initial population that
    already exceeds early sizes therefore saves the same state repeatedly until the match is popped."""
    r = ob.run(ob.make_opts(b0=1.0, b1=1.0, max_cells=200, initial={2: 51}, snapshots=[1, 11, 51, 61, 200]),
               hist_cap=64)
    assert r.n_snap_taken == 4  # 1, 11 and 51 popped at 51 cells, 61 at 61 cells; 200 is the final save
    np.testing.assert_array_equal(r.snap_cells[:4], [51, 51, 51, 61])
    assert np.all(r.snap_time[:3] == 0.0) and r.snap_time[3] > 0
    np.testing.assert_array_equal(r.snap_hist[0], r.snap_hist[2])


def _finals(n, **kw):
    out = ob.run_batch(ob.make_opts(**kw), 260, n, hist_cap=256)
    return out


def test_layouts_are_distribution_equivalent():
    """Reference layout + ChaCha8/ziggurat/BTPE vs GPU layout + Philox/popcount: same law.
    Two-sample KS on per-replicate summaries, p > 0.01 (BASELINE north star), 4000 replicates each."""
    n = 4000
    kw = dict(b0=1.0, b1=1.3, d0=0.1, d1=0.15, max_cells=1500)
    a = _finals(n, state=ob.STATE_VECTOR, rng=ob.RNG_RAND, **kw)
    b = _finals(n, state=ob.STATE_HIST, rng=ob.RNG_PHILOX, **kw)
    c = _finals(n, state=ob.STATE_VECTOR, rng=ob.RNG_PHILOX, **kw)
    alive = lambda r: r.stop == ob.STOP_MAX_CELLS
    # extinction probability agrees (binomial CI)
    pa, pb = 1 - alive(a).mean(), 1 - alive(b).mean()
    assert abs(pa - pb) < 4 * np.sqrt(pa * (1 - pa) * 2 / n) + 1e-3
    k = np.arange(256)
    for other in (b, c):
        for f in (lambda r: r.nplus[alive(r)].astype(float), lambda r: r.time[alive(r)].astype(float),
                  lambda r: (r.hist[alive(r)] * k).sum(axis=1) / r.hist[alive(r)].sum(axis=1)):
            assert sps.ks_2samp(f(a), f(other)).pvalue > 0.01
        # pooled final ecDNA distribution: chi-square homogeneity on well-filled classes
        ha, hb = a.hist[alive(a)].sum(axis=0).astype(float), other.hist[alive(other)].sum(axis=0).astype(float)
        keep = (ha + hb) > 4000
        # cells within a replicate are correlated, so compare class frequencies with a tolerance
        # derived from the between-replicate spread instead of a multinomial chi-square
        fa, fb = ha[keep] / ha.sum(), hb[keep] / hb.sum()
        assert np.max(np.abs(fa - fb)) < 0.01


def test_neutral_mean_is_a_martingale():
    """Neutral model from {1:1}: E[copies per cell over all cells] stays 1."""
    r = _finals(3000, state=ob.STATE_HIST, rng=ob.RNG_PHILOX, b0=1.0, b1=1.0, max_cells=800)
    mean = (r.hist * np.arange(256)).sum(axis=1) / r.hist.sum(axis=1)
    assert abs(mean.mean() - 1.0) < 4 * mean.std() / np.sqrt(len(mean))


def test_extinction_probability_matches_branching_theory():
    """No ecDNA advantage, b=1, d=0.4: a single-cell lineage dies out with probability d/b."""
    n = 6000
    r = _finals(n, state=ob.STATE_HIST, rng=ob.RNG_PHILOX, b0=1.0, b1=1.0, d0=0.4, d1=0.4, max_cells=300,
                max_time=1e6)
    p = (r.stop == ob.STOP_NO_INDIVIDUALS).mean()
    assert abs(p - 0.4) < 4 * np.sqrt(0.4 * 0.6 / n)


def test_event_type_frequencies_and_waiting_time():
    """First-reaction method == direct method: event i w.p. lambda_i / Lambda, dt ~ Exp(Lambda)."""
    init = {0: 300, 2: 200}
    rates = (1.0, 1.5, 0.3, 0.7)
    lam = np.array([rates[0] * 300, rates[1] * 200, rates[2] * 300, rates[3] * 200])
    ev, dts = [], []
    for i in range(6000):
        r = ob.run(ob.make_opts(b0=rates[0], b1=rates[1], d0=rates[2], d1=rates[3], max_cells=10 ** 6, max_iter=3,
                                initial=init, run_idx=i, max_time=100.0), hist_cap=64, trace_cap=4)
        ev.append(int(r.trace["event"][0]))
        dts.append(float(r.trace["dt"][0]))
    obs = np.bincount(ev, minlength=4)
    assert sps.chisquare(obs, lam / lam.sum() * len(ev)).pvalue > 1e-3
    assert sps.kstest(np.array(dts) * lam.sum(), "expon").pvalue > 1e-3


def test_subsample_is_an_exact_multivariate_hypergeometric_draw():
    """into_subsampled (main.rs:110-123) as the native mode defines it: sizes, bounds, the complement rule,
    determinism, and class frequencies that follow the population's."""
    hist = np.array([500, 0, 300, 150, 0, 40, 10], dtype=np.uint64)
    total = int(hist.sum())
    for want in (0, 1, 17, 499, 500, 501, 999, total, total + 5):
        out = ob.subsample(hist, want, seed=26, run_idx=260, j=0)
        assert int(out.sum()) == min(want, total)
        assert np.all(out <= hist)
        np.testing.assert_array_equal(out, ob.subsample(hist, want, seed=26, run_idx=260, j=0))
    assert np.any(ob.subsample(hist, 100, 26, 260, 0) != ob.subsample(hist, 100, 26, 261, 0))
    assert np.any(ob.subsample(hist, 100, 26, 260, 0) != ob.subsample(hist, 100, 26, 260, 1))
    # mean class counts over many replicates: E[out_k] = want * hist_k / total (both sides of the complement rule)
    for want in (200, 800):
        acc = np.zeros(len(hist))
        runs = 400
        for r in range(runs):
            acc += ob.subsample(hist, want, 7, 1000 + r, 3)
        expect = want * hist / total
        sd = np.sqrt(np.maximum(expect * (1 - hist / total), 1e-9) / runs)
        assert np.all(np.abs(acc / runs - expect) <= 5 * sd + 1e-9)
