"""Parity at BASELINE.json's FULL sizes, through ecdna_b200_run with the tile width the library's own planner
picks for the batch: replicates of the real C2 / C3 / C4 / C5 batches against the CPU oracle bit for bit, and the
north-star criterion for native mode (two-sample KS p > 0.01 and 99 % confidence intervals over 10^4 replicates)
at C1 and C3 size against the reference-layout oracle (per-cell vector, ChaCha8, ziggurat, BINV/BTPE).

Matches the reference's per-replicate closure, main.rs:92-99 / 166-173."""
import os

import numpy as np
import pytest

import oracle_binding as ob
from test_gpu_parity import WANT, assert_run_equal, assert_stats_equal, oracle_opts

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _check(pkg, res, o, picks, stride, **orc_kw):
    for i in picks:
        ref = ob.run(oracle_opts(o, o.idx_begin + i, **orc_kw), hist_cap=stride)
        assert_run_equal(res, i, ref, stride, digest=False)
        assert_stats_equal(res, i, ref.hist)
        yield i, ref


def test_c2_full_size_bit_exact(pkg, ctx):
    """C2: b1 = 1.5, 1 cell with 1 copy -> 1e6 cells, the whole batch of 1e4 replicates (1-lane tiles)."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.5, cells=1_000_000, runs=10_000, save_snapshots=False)
    res = ctx.run(o, want=WANT)
    assert res.timing.tile_width == pkg.plan(10_000)[0] == 1
    assert np.all(res.stop == pkg.STOP_MAX_CELLS) and np.all(res.n_events == 999_999)
    np.testing.assert_array_equal(res.hist.sum(axis=1), 1_000_000)
    assert res.timing.n_finished == 10_000
    list(_check(pkg, res, o, (0, 1, 4999, 7777, 9998, 9999), 512))


def test_c3_full_size_bit_exact_with_dynamics(pkg, ctx):
    """C3: birth-death d0 = d1 = 0.3, b1 = 1.2, 1e5 cells, 1e4 replicates, 300-point dynamics."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, d0=0.3, d1=0.3, cells=100_000, runs=10_000, save_snapshots=False)
    res = ctx.run(o, want=WANT + ("dyn", "dyn_count"), dyn_points=300, dyn_dt=0.1)
    grown = np.nonzero(res.stop == pkg.STOP_MAX_CELLS)[0]
    died = np.nonzero(res.stop == pkg.STOP_NO_INDIVIDUALS)[0]
    assert len(grown) > 6000 and len(died) > 1000
    picks = list(grown[:3]) + list(grown[-2:]) + list(died[:2]) + [int(np.argmax(res.n_events))]
    for i, ref in _check(pkg, res, o, picks, 512, dyn_points=300, dyn_dt=0.1):
        n = ref.dyn_count
        assert int(res.dyn_count[i]) == n
        np.testing.assert_array_equal(res.dyn[i][:n].view(np.uint32), ref.dyn[:n].view(np.uint32))


def test_c4_full_size_abc_distances(pkg, ctx):
    """C4: 1e5-cell birth-death runs with per-draw rates b1~U(1,2), d0,d1~U(0,0.5) and the fused ABC epilogue;
    65 536 draws of the 1e6 (same launch configuration: 1-lane tiles, full occupancy), 48 checked."""
    n = 65_536
    o = pkg.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=100_000, runs=n, save_snapshots=False)
    target = ob.run(oracle_opts(o, 260), hist_cap=512).hist
    np.testing.assert_array_equal(ctx.run(o, n_runs=1, idx_begin=260, want=("hist",)).hist[0].astype(np.uint64), target)
    rates = ctx.abc_draw_priors(seed=26, idx_begin=o.idx_begin, n_runs=n)
    thr = (0.05, 0.1, 0.1, 0.1)
    res = ctx.run(o, want=WANT + ("abc_distance", "abc_accept"), rates_per_run=rates, abc_target=target, abc_thresholds=thr)
    assert res.timing.tile_width == 1 and 0 < int(res.abc_accept.sum()) < n
    m = 48
    ref = ob.abc_batch(oracle_opts(o, 0), o.idx_begin, m, rates[:m], target, thr, THREADS, hist_cap=512)
    np.testing.assert_array_equal(res.n_events[:m], ref.n_events)
    np.testing.assert_array_equal(res.stop[:m], ref.stop)
    np.testing.assert_array_equal(res.abc_distance[:m].view(np.uint32), ref.distance.view(np.uint32))
    np.testing.assert_array_equal(res.abc_accept[:m], ref.accept)
    for i in (0, 17, 47):  # and the full state of three of them
        r1 = ob.run(oracle_opts(o, o.idx_begin + i, rates=rates[i]), hist_cap=512)
        assert_run_equal(res, i, r1, 512, digest=False)


def test_c5_full_size_bit_exact(pkg, ctx):
    """C5: {50: 1} -> 1e7 cells, neutral; the batch of 1e3 replicates with the default state mode (512-bin
    shared window, parking to the HBM arena when a replicate outgrows it), two replicates checked."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.0, cells=10_000_000, runs=1000, initial={50: 1}, save_snapshots=False)
    res = ctx.run(o, want=WANT)
    assert np.all(res.stop == pkg.STOP_MAX_CELLS) and np.all(res.n_events == 9_999_999)
    list(_check(pkg, res, o, (0, 999), 512))


def test_c5_hbm_resident_state_at_size(pkg, ctx):
    """The HBM-resident histogram (BASELINE config 5's state path) at 1e6 cells: same bits as the oracle."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.0, cells=1_000_000, runs=64, initial={50: 1}, save_snapshots=False)
    res = ctx.run(o, want=WANT, state_mode=pkg.STATE_HBM)
    list(_check(pkg, res, o, (0, 63), 512))


def _ks_and_ci(name, a, b):
    from scipy import stats as sps
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert sps.ks_2samp(a, b).pvalue > 0.01, name
    se = np.sqrt(a.var() / len(a) + b.var() / len(b))
    assert abs(a.mean() - b.mean()) < 2.58 * se + 1e-12, name


@pytest.mark.parametrize("name", ["C1", "C3"])
def test_native_matches_reference_layout_at_size(pkg, ctx, name):
    """North-star criterion for native mode at BASELINE size: 10^4 native replicates (Philox, histogram state,
    direct method) against 10^4 reference-layout replicates (ChaCha8, per-cell vector, first-reaction method):
    two-sample KS p > 0.01 on mean / frequency / entropy / clock and on the final ecDNA distribution, the means
    within 99 % confidence intervals, the same extinction probability."""
    from scipy import stats as sps
    kw = dict(C1=dict(b0=1.0, b1=1.0, cells=100_000), C3=dict(b0=1.0, b1=1.2, d0=0.3, d1=0.3, cells=100_000))[name]
    n = 10_000
    o = pkg.SimulationOptions(runs=n, save_snapshots=False, **kw)
    g = ctx.run(o, want=WANT, hist_stride=256)
    ref = ob.run_batch(oracle_opts(o, 0, state=ob.STATE_VECTOR, rng=ob.RNG_RAND), 10 ** 7, n, THREADS, hist_cap=256)
    ga, ra = g.stop == pkg.STOP_MAX_CELLS, ref.stop == ob.STOP_MAX_CELLS
    pg, pr = 1 - ga.mean(), 1 - ra.mean()
    assert abs(pg - pr) <= 2.58 * np.sqrt((pg * (1 - pg) + pr * (1 - pr)) / n) + 1e-12
    rstats = np.array([ob.stats(h) for h in ref.hist[ra]])
    _ks_and_ci("mean", g.mean[ga], rstats[:, 0])
    _ks_and_ci("frequency", g.frequency[ga], rstats[:, 1])
    _ks_and_ci("entropy", g.entropy[ga], rstats[:, 2])
    _ks_and_ci("clock", g.time[ga], ref.time[ra])
    # final ecDNA distribution: one random cell per replicate (independent samples), two-sample KS
    hg, hr = g.hist[ga].astype(np.float64), ref.hist[ra].astype(np.float64)
    fg, fr = hg / hg.sum(axis=1, keepdims=True), hr / hr.sum(axis=1, keepdims=True)
    rng = np.random.default_rng(5)

    def one_cell(f):
        u = rng.random(len(f))
        return (f.cumsum(axis=1) < u[:, None]).sum(axis=1)

    assert sps.ks_2samp(one_cell(fg), one_cell(fr)).pvalue > 0.01
    # pooled distribution: per-class frequencies within the between-replicate spread
    se = np.sqrt(fg.var(axis=0) / len(fg) + fr.var(axis=0) / len(fr)) + 1e-9
    z = np.abs(fg.mean(axis=0) - fr.mean(axis=0)) / se
    busy = (fg.mean(axis=0) + fr.mean(axis=0)) > 1e-3
    assert (z[busy] > 3.5).sum() <= 1 and z[busy].max() < 5.0


def test_native_stream_exact_moments_at_1e6_replicates(pkg, ctx):
    """Known answers of the process itself (no crate semantics involved), at a statistical power only the GPU
    affords: 10^6 replicates each.
      * Yule process (pure birth, rate b per cell) from 1 to N cells: the clock at the stop is a sum of independent
        Exp(i b) waiting times, E[T] = sum 1/(i b), Var[T] = sum 1/(i b)^2  -> the direct method's waiting times;
      * neutral growth: the mean copy number over all cells is a martingale, E = the initial mean (1.0)
        -> the cell pick and the Binomial(2k, 1/2) split;
      * linear birth-death from one cell with b = 1, d = 1/2: P(extinction before N cells) =
        (q - q^N) / (1 - q^N), q = d / b -> the choice of the reaction."""
    n = 1_000_000
    # Yule: only ecDNA- cells (no segregation involved): initial {0: 1}
    N, b = 300, 1.5
    o = pkg.SimulationOptions(b0=b, b1=1.0, runs=n, initial={0: 1}, years=10_000, save_snapshots=False)  # (no time stop)
    o.max_cells = N
    r = ctx.run(o, want=("stop_reason", "time", "n_events", "nminus"))
    assert np.all(r.stop == pkg.STOP_MAX_CELLS) and np.all(r.n_events == N - 1)
    i = np.arange(1, N, dtype=np.float64)
    mu, var = (1 / (i * b)).sum(), (1 / (i * b) ** 2).sum()
    t = r.time.astype(np.float64)
    assert abs(t.mean() - mu) < 5 * np.sqrt(var / n), (t.mean(), mu)
    assert abs(t.var() - var) < 6 * var * np.sqrt(2.5 / n), (t.var(), var)  # (kurtosis of a sum of exponentials < 9)
    # neutral martingale
    o = pkg.SimulationOptions(b0=1.0, b1=1.0, cells=400, runs=n, save_snapshots=False)
    r = ctx.run(o, want=("stop_reason", "mean", "nplus"))
    m = r.mean.astype(np.float64)
    assert abs(m.mean() - 1.0) < 5 * m.std() / np.sqrt(n), (m.mean(), m.std())
    # extinction probability of a linear birth-death process from one cell
    N, q = 60, 0.5
    o = pkg.SimulationOptions(b0=1.0, b1=1.0, d0=q, d1=q, runs=n, initial={0: 1}, years=10_000, save_snapshots=False)
    o.max_cells = N
    r = ctx.run(o, want=("stop_reason",))
    p_ext = (q - q ** N) / (1 - q ** N)
    got = float((r.stop == pkg.STOP_NO_INDIVIDUALS).mean())
    assert set(np.unique(r.stop)) <= {pkg.STOP_NO_INDIVIDUALS, pkg.STOP_MAX_CELLS}
    assert abs(got - p_ext) < 5 * np.sqrt(p_ext * (1 - p_ext) / n), (got, p_ext)
