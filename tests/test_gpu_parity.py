"""GPU parity: libecdna_b200.so (through its C ABI) against the CPU oracle, same seeded inputs.

Integer results (stop reason, counts, event numbers, histograms, digests) and the f32 clock must be
BIT-EXACT: the oracle's histogram/philox configuration is the specification of the kernel's native
mode, and replay mode consumes the decision stream of the reference-layout (vector, ChaCha8) oracle.
Summary statistics are exact too: the integer moments are exact, every float step is a single IEEE operation in a
fixed order, and the entropy is a sum of fixed-point terms (entropy_term_q40), which does not depend on the order.
"""
import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def assert_stats_equal(res, i, hist):
    """mean / frequency / entropy / variance of replicate i: the oracle's bits."""
    want = np.array(ob.stats(hist), dtype=np.float32)
    got = np.array([res.mean[i], res.frequency[i], res.entropy[i], res.variance[i]], dtype=np.float32)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32), err_msg=f"run {i}: {got} vs {want}")



def oracle_opts(o, run_idx, state=ob.STATE_HIST, rng=ob.RNG_PHILOX, rates=None, **kw):
    b0, b1, d0, d1 = rates if rates is not None else (o.b0, o.b1, o.d0, o.d1)
    return ob.make_opts(b0=b0, b1=b1, d0=d0, d1=d1, segregation=o.segregation, state=state, rng=rng,
                        max_cells=o.max_cells, max_iter=o.max_iter, max_time=float(o.years), seed=o.seed,
                        run_idx=run_idx, initial=o.distribution, **kw)


def assert_run_equal(res, i, ref, stride, digest=True):
    assert int(res.stop[i]) == ref.stop_reason, f"run {i}: stop {res.stop[i]} vs {ref.stop_reason}"
    assert int(res.n_events[i]) == ref.n_events
    assert int(res.nminus[i]) == ref.nminus and int(res.nplus[i]) == ref.nplus
    assert np.float32(res.time[i]).view(np.uint32) == np.float32(ref.time).view(np.uint32)
    assert int(res.kmax[i]) == ref.kmax
    np.testing.assert_array_equal(res.hist[i].astype(np.uint64), ref.hist[:stride])
    assert int(res.sum_k[i]) == ref.sum_k
    assert int(res.n_div[i]) == ref.n_div and int(res.n_death[i]) == ref.n_death
    if digest:
        assert int(res.hash[i]) == ref.hash
        assert int(res.chain[i]) == ref.chain


WANT = ("stop_reason", "nminus", "nplus", "time", "n_events", "kmax", "hist", "hash", "chain", "sum_k", "n_div",
        "n_death", "mean", "frequency", "entropy", "variance")

CASES = {
    "neutral": dict(b0=1.0, b1=1.0, cells=20000),
    "selection": dict(b0=1.0, b1=1.5, cells=20000),
    "birth_death": dict(b0=1.0, b1=1.2, d0=0.3, d1=0.3, cells=10000),
    "death_nplus_only": dict(b0=1.0, b1=1.1, d1=0.4, cells=5000),
    "deterministic": dict(b0=1.0, b1=1.3, cells=5000, segregation="deterministic", initial={4: 3, 0: 2}),
    "no_uneven": dict(b0=1.0, b1=1.0, cells=5000, segregation="binomial-no-uneven"),
    "no_nminus": dict(b0=0.9, b1=1.0, cells=5000, segregation="binomial-no-nminus"),
    "k50": dict(b0=1.0, b1=1.0, cells=20000, initial={50: 1}),
    "multi_bin_initial": dict(b0=1.0, b1=1.2, d0=0.1, d1=0.2, cells=8000, initial={0: 10, 2: 5, 33: 4, 64: 1, 7: 9}),
    "time_stop": dict(b0=1.0, b1=1.0, years=5),
    "extinction": dict(b0=0.5, b1=0.5, d0=1.0, d1=1.0, cells=1000, initial={3: 20, 0: 10}),
    # copy numbers beyond the smallest shared-memory window (128 bins): exercise parking / HBM state
    "wide": dict(b0=1.0, b1=1.3, cells=8000, initial={100: 1}),
    "wide_bd": dict(b0=1.0, b1=1.2, d0=0.1, d1=0.2, cells=6000, initial={120: 3, 0: 5, 7: 2}),
    "wide_initial": dict(b0=1.0, b1=1.1, cells=3000, initial={200: 2, 3: 1, 130: 1}),
    # copy numbers near the top of u16: segregation draws of up to 62000 bits (hundreds of Philox slots),
    # HBM arena walked over hundreds of rows, and the u16 doubling overflow somewhere along the way
    "huge_k": dict(b0=1.0, b1=1.0, d1=0.2, cells=40, initial={20000: 2, 31000: 1, 9: 1}),
}


@pytest.mark.parametrize("digest", [False, True], ids=["straight", "digest"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_native_bit_exact(pkg, ctx, name, digest):
    """Native (Philox) mode reproduces the histogram oracle bit for bit, every replicate.  Without the
    digest the kernel runs its straight-line event step (rare events redone by the complete step);
    with it every event takes the complete step and the per-event chain digest is compared too."""
    o = pkg.SimulationOptions(runs=12, save_snapshots=False, **CASES[name])
    stride = 65536 if name == "huge_k" else 512
    res = ctx.run(o, want=WANT, digest=digest, hist_stride=stride)
    for i in range(o.runs):
        ref = ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=stride)
        assert_run_equal(res, i, ref, stride, digest=digest)
        assert_stats_equal(res, i, ref.hist)
    assert res.timing.total_events == int(res.n_events.sum())


@pytest.mark.parametrize("tile_width", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("name", ["selection", "birth_death", "k50", "no_uneven", "multi_bin_initial", "extinction"])
def test_tile_widths_agree(pkg, ctx, name, tile_width):
    """Sub-warp tiles (several replicates per warp) give the same bits as one warp per replicate."""
    o = pkg.SimulationOptions(runs=37, save_snapshots=False, **CASES[name])
    a = ctx.run(o, want=WANT, digest=True, tile_width=32)
    for digest in (True, False):
        b = ctx.run(o, want=WANT, digest=digest, tile_width=tile_width)
        for f in ("stop_reason", "nminus", "nplus", "n_events", "kmax", "hist", "sum_k") + (("hash", "chain") if digest else ()):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
        np.testing.assert_array_equal(a.time.view(np.uint32), b.time.view(np.uint32))
        for f in ("mean", "frequency", "entropy", "variance"):  # (exact moments + fixed-point entropy: no order dependence)
            np.testing.assert_array_equal(getattr(a, f).view(np.uint32), getattr(b, f).view(np.uint32), err_msg=f)


@pytest.mark.parametrize("mode", ["hbm", "spill_resume", "spill_restart", "spill_mixed_l8", "lane_cascade", "lane_cascade_restart"])
@pytest.mark.parametrize("name", ["wide", "wide_bd", "wide_initial", "huge_k"])
def test_hbm_state_bit_exact(pkg, ctx, name, mode):
    """The HBM-resident histogram, and parking a replicate that outgrows shared memory (with its
    state saved, or restarted from event 0 when no record slot is left), do not change a single bit."""
    o = pkg.SimulationOptions(runs=10, save_snapshots=False, **CASES[name])
    kw = {"hbm": dict(state_mode=pkg.STATE_HBM), "spill_resume": dict(smem_bins=128),
          "spill_restart": dict(smem_bins=128, spill_records=0xFFFFFFFF),
          "spill_mixed_l8": dict(smem_bins=128, spill_records=3, tile_width=8),
          # 1-lane tiles with a 128-bin window: parked replicates continue in a 256-bin launch, its leftovers in HBM
          "lane_cascade": dict(smem_bins=128, tile_width=1),
          "lane_cascade_restart": dict(smem_bins=128, tile_width=1, spill_records=2)}[mode]
    stride = 65536 if name == "huge_k" else 512
    for digest in (True, False):
        res = ctx.run(o, want=WANT, digest=digest, hist_stride=stride, **kw)
        for i in range(o.runs):
            ref = ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=stride)
            assert_run_equal(res, i, ref, stride, digest=digest)
    if mode.startswith("lane"):
        assert res.timing.kernel_launches == 3 and res.timing.smem_bins == 128
    elif mode != "hbm":
        assert res.timing.n_spilled > 0 and np.any(res.stop_reason & pkg.FLAG_SPILLED)
        assert res.timing.kernel_launches == 2


SLICE_CASES = {
    # tile width, replicates (more than the tiles of one block per SM: 148 * 128 / width), case
    "warp_selection": (32, 700, dict(b0=1.0, b1=1.3, cells=600)),
    # early extinctions free tiles: whether anyone still waits at a slice boundary depends on timing
    "warp_bd": (32, 900, dict(b0=1.0, b1=1.2, d0=0.3, d1=0.3, cells=600)),
    "l4_bd": (4, 6000, dict(b0=1.0, b1=1.2, d0=0.3, d1=0.3, cells=400)),
    "l4_selection": (4, 5200, dict(b0=1.0, b1=1.5, cells=400)),
    "l8_no_uneven": (8, 2600, dict(b0=1.0, b1=1.0, cells=500, segregation="binomial-no-uneven")),
    "l16_k50": (16, 1300, dict(b0=1.0, b1=1.0, cells=500, initial={50: 1})),
    "l2_bd": (2, 10000, dict(b0=1.0, b1=1.3, d0=0.2, d1=0.1, cells=300)),
    "l2_selection": (2, 9600, dict(b0=1.0, b1=1.5, cells=300)),
}


@pytest.mark.parametrize("digest", [False, True], ids=["straight", "digest"])
@pytest.mark.parametrize("name", sorted(SLICE_CASES))
def test_time_slicing_bit_exact(pkg, ctx, name, digest):
    """A batch larger than the launch holds is time-sliced: every 64 events a replicate saves its state
    and makes room for the one that waits longest, and is resumed later by whichever tile is free.
    Every replicate still matches the oracle bit for bit (the draws are keyed by replicate and event)."""
    tw, runs, case = SLICE_CASES[name]
    o = pkg.SimulationOptions(runs=runs, save_snapshots=False, **case)
    res = ctx.run(o, want=WANT, digest=digest, tile_width=tw, slice_events=64)
    assert res.timing.slice_events == 64
    if "_bd" not in name:
        assert res.timing.n_slices > 0
    for i in range(o.runs):
        ref = ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=512)
        assert_run_equal(res, i, ref, 512, digest=digest)
    assert res.timing.total_events == int(res.n_events.sum())


def test_time_slicing_keeps_snapshots_dynamics_and_parking(pkg, ctx):
    """Snapshot cursor, dynamics cursor and the roofline counters travel with a sliced replicate, and a
    sliced replicate that outgrows the shared window still moves to the HBM launch."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, d0=0.2, d1=0.2, cells=1500, runs=5000, initial={2: 40, 0: 11},
                              snapshots=[1, 51, 60, 500, 1000, 1500])
    want = WANT + ("snap_count", "snap_cells", "snap_time", "snap_hist", "dyn", "dyn_count")
    kw = dict(want=want, tile_width=4, dyn_points=100, dyn_dt=0.1)
    a = ctx.run(o, slice_events=0xFFFFFFFF, **kw)
    b = ctx.run(o, slice_events=128, **kw)
    assert a.timing.slice_events == 0 and a.timing.n_slices == 0
    assert b.timing.slice_events == 128 and b.timing.n_slices > 0
    for f in want:
        x, y = getattr(a, f), getattr(b, f)
        if x.dtype.kind == "f":
            x, y = x.view(np.uint32), y.view(np.uint32)
        np.testing.assert_array_equal(x, y, err_msg=f)
    # parking: copy numbers beyond a 128-bin window
    o2 = pkg.SimulationOptions(runs=700, save_snapshots=False, b0=1.0, b1=1.3, cells=3000, initial={50: 1})
    c = ctx.run(o2, want=WANT, tile_width=32, smem_bins=128, slice_events=0xFFFFFFFF)
    d = ctx.run(o2, want=WANT, tile_width=32, smem_bins=128, slice_events=64)
    assert d.timing.n_slices > 0 and d.timing.n_spilled > 0 and d.timing.kernel_launches == 2
    for f in ("stop_reason", "nminus", "nplus", "n_events", "kmax", "hist", "sum_k", "n_div", "n_death"):
        np.testing.assert_array_equal(getattr(c, f), getattr(d, f), err_msg=f)
    np.testing.assert_array_equal(c.time.view(np.uint32), d.time.view(np.uint32))


def test_automatic_slicing_only_where_it_pays(pkg, ctx):
    """slice_events = 0: small batches and batches of many waves run as before; a batch a little larger
    than two blocks per SM hold is sliced."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.5, cells=300, save_snapshots=False)
    small = ctx.run(o, n_runs=64, want=("n_events",), tile_width=4)
    assert small.timing.slice_events == 0
    mid = ctx.run(o, n_runs=10000, want=("n_events",), tile_width=4)
    assert mid.timing.slice_events > 0 and mid.timing.grid_blocks == 2 * 148
    ref = ctx.run(o, n_runs=10000, want=("n_events",), tile_width=4, slice_events=0xFFFFFFFF)
    assert ref.timing.slice_events == 0 and ref.timing.grid_blocks == 313
    np.testing.assert_array_equal(mid.n_events, ref.n_events)


def test_randomized_differential_across_tile_widths_and_slicing(pkg, ctx):
    """Twelve random parameter sets (rates with zeros and extremes, all segregation rules, sparse initial
    distributions, stop by size or by time): every tile width, with and without time slicing, gives the bits
    of the one-warp-per-replicate digest run, and a sample of replicates equals the oracle."""
    rng = np.random.default_rng(20261018)
    rules = ["binomial", "deterministic", "binomial-no-uneven", "binomial-no-nminus"]
    for case in range(12):
        b0, b1 = rng.choice([0.0, 0.5, 1.0, 1.7]), rng.choice([0.8, 1.0, 1.5, 3.0])
        d0, d1 = rng.choice([0.0, 0.0, 0.2, 0.9]), rng.choice([0.0, 0.0, 0.3, 1.1])
        init = {int(k): int(c) for k, c in zip(rng.choice(np.arange(1, 70), size=3, replace=False), rng.integers(1, 6, size=3))}
        if rng.random() < 0.5:
            init[0] = int(rng.integers(1, 8))
        kw = dict(b0=float(b0), b1=float(b1), d0=float(d0), d1=float(d1), cells=int(rng.integers(60, 2500)),
                  segregation=str(rng.choice(rules)), initial=init, seed=int(rng.integers(1, 1000)))
        if rng.random() < 0.3:  # stop by time instead of size (clap group 'stop')
            del kw["cells"]
            kw["years"] = int(rng.integers(1, 6))
        o = pkg.SimulationOptions(runs=700, save_snapshots=False, **kw)
        ref = ctx.run(o, want=WANT, digest=True, tile_width=32, slice_events=0xFFFFFFFF)
        for i in range(0, o.runs, 97):
            assert_run_equal(ref, i, ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=512), 512, digest=True)
        for tw, sl in ((1, 0), (2, 0xFFFFFFFF), (4, 0xFFFFFFFF), (8, 64), (16, 0xFFFFFFFF), (32, 64), (2, 0)):
            got = ctx.run(o, want=WANT, digest=False, tile_width=tw, slice_events=sl)
            for f in ("stop_reason", "nminus", "nplus", "n_events", "kmax", "hist", "sum_k", "n_div", "n_death"):
                np.testing.assert_array_equal(getattr(ref, f), getattr(got, f), err_msg=f"case {case} {kw} tile {tw} {f}")
            np.testing.assert_array_equal(ref.time.view(np.uint32), got.time.view(np.uint32), err_msg=f"case {case} tile {tw}")


def test_smem_only_mode_reports_overflow(pkg, ctx):
    """state_mode SMEM never parks: a replicate that outgrows the window stops with HIST_OVERFLOW."""
    o = pkg.SimulationOptions(runs=4, save_snapshots=False, **CASES["wide"])
    res = ctx.run(o, want=WANT, state_mode=pkg.STATE_SMEM, smem_bins=128)
    assert np.all(res.stop == pkg.STOP_HIST_OVERFLOW) and res.timing.kernel_launches == 1


def _vector_trace(o, run_idx, rng):
    opts = oracle_opts(o, run_idx, state=ob.STATE_VECTOR, rng=rng)
    cap = int(o.max_cells) * 8
    return ob.run(opts, hist_cap=512, trace_cap=cap)


@pytest.mark.parametrize("rng", [ob.RNG_RAND, ob.RNG_PHILOX])
@pytest.mark.parametrize("name", ["neutral", "selection", "birth_death", "no_nminus", "k50", "extinction"])
def test_replay_bit_exact(pkg, ctx, name, rng):
    """Replay mode: the decision stream of the reference-layout oracle (per-cell vector, swap-remove,
    ChaCha8 + rand conversions, or Philox) drives the kernel to the same state after every event."""
    o = pkg.SimulationOptions(runs=6, save_snapshots=False, **CASES[name])
    traces, refs = [], []
    for i in range(o.runs):
        r = _vector_trace(o, o.idx_begin + i, rng)
        assert r.trace_len == r.n_events == len(r.trace)
        traces.append(r.trace)
        refs.append(r)
    offsets = np.zeros(o.runs + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(t) for t in traces])
    stream = np.concatenate(traces) if offsets[-1] else np.zeros(1, dtype=ob.REPLAY_DTYPE)
    res = ctx.run(o, want=WANT, digest=True, replay=stream, replay_offsets=offsets)
    for i, ref in enumerate(refs):
        # the stream ends exactly where the reference's stop rule fires, so the stop reason matches too
        assert_run_equal(res, i, ref, 512)
        # and the histogram oracle replaying the same stream agrees as well
        h = ob.run(oracle_opts(o, o.idx_begin + i, rng=ob.RNG_REPLAY, replay=traces[i]), hist_cap=512)
        assert (h.hash, h.chain, h.n_events) == (ref.hash, ref.chain, ref.n_events)


def test_kernel_matches_golden_fixtures(pkg, ctx):
    """The committed fixtures (tests/golden/golden_v2.json, frozen from the oracle) through the C ABI."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.json")))
    seg_names = {v: k for k, v in pkg.SEGREGATION_NAMES.items()}
    for name, kw in make_golden.CASES.items():
        o = pkg.SimulationOptions(b0=kw.get("b0", 1.0), b1=kw.get("b1", 1.0), d0=kw.get("d0"), d1=kw.get("d1"),
                                  cells=kw["max_cells"], initial=kw.get("initial"), runs=4, save_snapshots=False,
                                  segregation=seg_names[kw.get("segregation", ob.SEG_BINOMIAL)])
        for tile in (0, 32, 4, 1):
            res = ctx.run(o, want=WANT, digest=tile != 1, tile_width=tile)
            for i, want in enumerate(gold["native"][name]):
                got = {"stop": int(res.stop[i]), "nminus": int(res.nminus[i]), "nplus": int(res.nplus[i]),
                       "n_events": int(res.n_events[i]), "time_bits": int(res.time[i:i + 1].view(np.uint32)[0]),
                       "kmax": int(res.kmax[i]), "hash": int(res.hash[i]), "chain": int(res.chain[i]),
                       "hist": {str(k): int(c) for k, c in enumerate(res.hist[i]) if c}}
                if tile == 1:  # the straight-line step keeps no digest
                    got["hash"], got["chain"] = want["hash"], want["chain"]
                assert got == want, (name, tile, i)


def test_cli_writes_reference_layout(pkg, ctx, tmp_path):
    """The drop-in `ecdna` binary: same flags, same files (process.rs:39-45, lib.rs:27-45) and the
    histograms in them equal the library's (and therefore the oracle's)."""
    import json
    import os
    import subprocess
    pkg.build()
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "host", "ecdna")
    out = str(tmp_path / "out")
    cmd = [exe, "--b1", "1.5", "--d0", "0.1", "--d1", "0.2", "--cells", "500", "--runs", "3", "--seed", "7",
           "--snapshots=1,100,500", "--subsamples=50", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Starting the simulation" in r.stdout and "End simulation" in r.stdout
    o = pkg.SimulationOptions(b1=1.5, d0=0.1, d1=0.2, cells=500, runs=3, seed=7, snapshots=[1, 100, 500],
                              subsamples=[50])
    res = ctx.run(o, want=WANT + ("snap_count", "snap_cells", "snap_time", "snap_hist", "sub_hist"))
    for i in range(3):
        name = f"1b0_1dot5b1_0dot1d0_0dot2d1_{70 + i}idx.json"
        cells = int(res.nminus[i] + res.nplus[i])
        t = f"{float(res.time[i]):.1f}".replace(".", "dot") + "years"
        final = json.load(open(os.path.join(out, f"{cells}cells", "ecdna", t, name)))
        assert final == {str(k): int(c) for k, c in enumerate(res.hist[i]) if c}
        for s in range(int(res.snap_count[i])):
            sc = int(res.snap_cells[i][s])
            st = f"{float(res.snap_time[i][s]):.1f}".replace(".", "dot") + "years"
            snap = json.load(open(os.path.join(out, f"{sc}cells", "ecdna", st, name)))
            assert sum(snap.values()) == sc
        if cells > 50:
            sub = [p for p in os.listdir(os.path.join(out, "50cells", "ecdna")) if p == t]
            assert sub, "subsample directory missing"
            got = json.load(open(os.path.join(out, "50cells", "ecdna", t, name)))
            assert sum(got.values()) == 50
            assert got == {str(k): int(c) for k, c in enumerate(res.sub_hist[i, 0]) if c}
    # the optional outputs of the pre-0.19 reference: summaries and dynamics
    out2 = str(tmp_path / "out2")
    r = subprocess.run([exe, "--b1", "1.2", "--cells", "300", "--runs", "2", "--summaries", "--dynamics", out2],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o2 = pkg.SimulationOptions(b1=1.2, cells=300, runs=2)
    res2 = ctx.run(o2, want=WANT + ("dyn", "dyn_count"), dyn_points=300, dyn_dt=0.1)
    for i in range(2):
        t2 = f"{float(res2.time[i]):.1f}".replace(".", "dot") + "years"
        name2 = f"1b0_1dot2b1_0d0_0d1_{260 + i}idx.json"
        for what, val in (("mean", res2.mean[i]), ("frequency", res2.frequency[i]), ("entropy", res2.entropy[i])):
            got = float(open(os.path.join(out2, "300cells", what, t2, name2)).read())
            assert np.float32(got) == np.float32(val), what
        dyn = json.load(open(os.path.join(out2, "300cells", "dynamics", t2, name2)))
        n = int(res2.dyn_count[i])
        assert len(dyn["nplus"]) == n and dyn["dt"] == 0.1
        np.testing.assert_array_equal(np.array(dyn["nplus"], dtype=np.float32), res2.dyn[i][:n, 1])
        np.testing.assert_array_equal(np.array(dyn["mean"], dtype=np.float32), res2.dyn[i][:n, 2])
    # clap-compatible errors
    assert subprocess.run([exe, "--years", "3", "--cells", "4", out], capture_output=True).returncode == 2
    assert subprocess.run([exe], capture_output=True).returncode == 2


@pytest.mark.parametrize("name", ["neutral", "selection", "birth_death", "no_uneven", "no_nminus", "k50",
                                  "deterministic", "multi_bin_initial", "extinction", "time_stop"])
def test_uniform_level_replay_bit_exact(pkg, ctx, name):
    """rng_mode UNIFORMS: the kernel consumes the raw u64 stream of the reference-layout oracle's
    ChaCha8 generator and re-runs the replicate on the reference's own per-cell layout.  Same state
    after every event (chain digest), same final distribution, clock bits and stop reason."""
    kw = dict(CASES[name])
    if name == "time_stop":
        kw = dict(b0=1.0, b1=1.0, years=4)
    o = pkg.SimulationOptions(runs=5, save_snapshots=False, **kw)
    streams, refs = [], []
    for i in range(o.runs):
        r = ob.run(oracle_opts(o, o.idx_begin + i, state=ob.STATE_VECTOR, rng=ob.RNG_RAND), hist_cap=512,
                   u64_cap=4_000_000)
        assert r.u64_len == len(r.u64)
        streams.append(r.u64)
        refs.append(r)
    offsets = np.zeros(o.runs + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in streams])
    stream = np.concatenate(streams) if offsets[-1] else np.zeros(1, dtype=np.uint64)
    if o.max_cells > 10 ** 6:  # the per-cell arena is sized by max_cells
        o.max_cells = int(max(r.nminus + r.nplus for r in refs)) + 8
    res = ctx.run(o, want=WANT[:12], replay_u64=stream, replay_offsets=offsets)
    for i, ref in enumerate(refs):
        assert_run_equal(res, i, ref, 512)
    # a stream that ends early is reported, not read past
    cut = ctx.run(o, n_runs=1, want=WANT[:12], replay_u64=streams[0][: len(streams[0]) // 2],
                  replay_offsets=np.array([0, len(streams[0]) // 2], dtype=np.uint64))
    if len(streams[0]) > 4:
        assert int(cut.stop[0]) == pkg.STOP_REPLAY_END and int(cut.n_events[0]) < refs[0].n_events


def test_uniform_replay_overflow_reports_the_reference_state(pkg, ctx):
    """k >= 32768 cannot double (proliferation.rs:63-67 panics with the cell already taken out at :57): the
    uniform-level kernel, the histogram kernel and the oracle all stop there with the same state."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.0, cells=60, runs=3, initial={40000: 1, 3: 2}, save_snapshots=False)
    streams, refs = [], []
    for i in range(o.runs):
        r = ob.run(oracle_opts(o, o.idx_begin + i, state=ob.STATE_VECTOR, rng=ob.RNG_RAND), hist_cap=65536, u64_cap=100_000)
        streams.append(r.u64)
        refs.append(r)
    assert any(r.stop_reason == ob.STOP_COPY_OVERFLOW for r in refs)
    offsets = np.zeros(o.runs + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in streams])
    res = ctx.run(o, want=WANT[:12], replay_u64=np.concatenate(streams), replay_offsets=offsets, hist_stride=65536)
    for i, ref in enumerate(refs):
        assert_run_equal(res, i, ref, 65536)
    # the native stream stops the same way (its own draws: only the stop code and the bookkeeping are compared)
    nat = ctx.run(o, want=WANT, hist_stride=65536, digest=True)
    for i in range(o.runs):
        assert_run_equal(nat, i, ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=65536), 65536)


def test_replay_detects_inconsistency(pkg, ctx):
    o = pkg.SimulationOptions(runs=1, cells=200, save_snapshots=False)
    r = _vector_trace(o, o.idx_begin, ob.RNG_RAND)
    bad = r.trace.copy()
    j = int(np.nonzero(bad["event"] == ob.EV_BIRTH_NPLUS)[0][3])
    bad["k"][j] = 400  # no such cell
    offsets = np.array([0, len(bad)], dtype=np.uint64)
    res = ctx.run(o, want=WANT, replay=bad, replay_offsets=offsets)
    assert int(res.stop[0]) == pkg.STOP_REPLAY_BAD and int(res.n_events[0]) == j
    short = r.trace[:50].copy()
    res = ctx.run(o, want=WANT, replay=short, replay_offsets=np.array([0, 50], dtype=np.uint64))
    assert int(res.stop[0]) == pkg.STOP_REPLAY_END and int(res.n_events[0]) == 50


def test_subsamples_match_oracle(pkg, ctx):
    """main.rs:110-123: the --subsamples draws on the device, one warp per (replicate, size), bit for bit the
    oracle's exact multivariate hypergeometric draw from the final distribution."""
    sizes = [0, 1, 50, 700, 1400, 2999, 3000, 5000]
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, d0=0.2, d1=0.2, cells=3000, runs=9, initial={2: 40, 0: 11},
                              save_snapshots=False, subsamples=sizes)
    res = ctx.run(o, want=WANT + ("sub_hist",), hist_stride=256)
    assert res.sub_hist.shape == (o.runs, len(sizes), 256)
    for i in range(o.runs):
        final = res.hist[i].astype(np.uint64)
        for j, n in enumerate(sizes):
            ref = ob.subsample(final, n, o.seed, o.idx_begin + i, j)
            np.testing.assert_array_equal(res.sub_hist[i, j].astype(np.uint64), ref, err_msg=f"run {i} size {n}")
            assert int(res.sub_hist[i, j].sum()) == min(n, int(final.sum()))
    # without the final distribution among the outputs the library keeps one on the device for itself
    only = ctx.run(o, want=("sub_hist",), hist_stride=256)
    np.testing.assert_array_equal(only.sub_hist, res.sub_hist)


def test_snapshots_match_oracle(pkg, ctx):
    """process.rs:122-145: snapshot capture against the pre-event population, front-pop quirk included."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, d0=0.2, d1=0.2, cells=3000, runs=8, initial={2: 40, 0: 11},
                              snapshots=[1, 51, 60, 500, 1000, 3000])
    want = WANT + ("snap_count", "snap_cells", "snap_time", "snap_hist")
    res = ctx.run(o, want=want, digest=False, tile_width=4)
    for i in range(o.runs):
        ref = ob.run(oracle_opts(o, o.idx_begin + i, snapshots=o.snapshots), hist_cap=512)
        assert_run_equal(res, i, ref, 512, digest=False)
        assert int(res.snap_count[i]) == ref.n_snap_taken
        n = ref.n_snap_taken
        np.testing.assert_array_equal(res.snap_cells[i][:n], ref.snap_cells[:n])
        np.testing.assert_array_equal(res.snap_time[i][:n].view(np.uint32), ref.snap_time[:n].view(np.uint32))
        np.testing.assert_array_equal(res.snap_hist[i][:n].astype(np.uint64), ref.snap_hist[:n])


def test_default_snapshots_pure_birth(pkg, ctx):
    o = pkg.SimulationOptions(cells=2000, runs=5)  # 11 default sizes, clap_app.rs:121-134
    assert o.snapshots == [1, 201, 401, 601, 801, 1001, 1201, 1401, 1601, 1801, 2000]
    res = ctx.run(o, want=WANT + ("snap_count", "snap_cells", "snap_hist"))
    # the last size equals max_cells: the run stops before an event sees it (the final save covers it)
    assert np.all(res.snap_count == 10)
    np.testing.assert_array_equal(res.snap_hist.sum(axis=2)[:, :10], res.snap_cells[:, :10])


def test_dynamics_match_oracle(pkg, ctx):
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, d0=0.3, d1=0.3, cells=5000, runs=6, save_snapshots=False)
    res = ctx.run(o, want=WANT + ("dyn", "dyn_count"), dyn_points=300, dyn_dt=0.1)
    for i in range(o.runs):
        ref = ob.run(oracle_opts(o, o.idx_begin + i, dyn_points=300, dyn_dt=0.1), hist_cap=512)
        assert int(res.dyn_count[i]) == ref.dyn_count
        n = ref.dyn_count
        np.testing.assert_array_equal(res.dyn[i][:n].view(np.uint32), ref.dyn[:n].view(np.uint32))


def test_abc_epilogue_matches_oracle(pkg, ctx):
    """abc.md:38-55: per-draw distances to the target distribution, fused into the kernel's tail."""
    n = 64
    o = pkg.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=3000, runs=n, save_snapshots=False)
    target = ob.run(oracle_opts(o, 260), hist_cap=512).hist
    rates = ctx.abc_draw_priors(seed=26, idx_begin=o.idx_begin, n_runs=n)
    assert np.all(rates[:, 0] == 1.0) and np.all((rates[:, 1] >= 1.0) & (rates[:, 1] <= 2.0))
    thr = (0.2, 0.5, 0.5, 0.5)
    res = ctx.run(o, want=WANT + ("abc_distance", "abc_accept"), rates_per_run=rates, abc_target=target,
                  abc_thresholds=thr, digest=False, tile_width=4)
    n_acc = 0
    for i in range(n):
        ref = ob.run(oracle_opts(o, o.idx_begin + i, rates=rates[i]), hist_cap=512)
        assert_run_equal(res, i, ref, 512, digest=False)
        assert_stats_equal(res, i, ref.hist)
        n_acc += int(res.abc_accept[i])
    assert 0 < n_acc < n
    # the four distances and the accept flag: the oracle's bits, every draw
    batch = ob.abc_batch(oracle_opts(o, 0), o.idx_begin, n, rates, target, thr, 0, hist_cap=512)
    np.testing.assert_array_equal(res.abc_distance.view(np.uint32), batch.distance.view(np.uint32))
    np.testing.assert_array_equal(res.abc_accept, batch.accept)


def test_native_distribution_matches_reference_layout(pkg, ctx):
    """North-star criterion for native mode: against the reference-layout oracle (per-cell vector,
    ChaCha8 + rand conversions, ziggurat, BINV/BTPE) over 10^4 replicates each, two-sample KS p > 0.01
    on the per-replicate mean / frequency / entropy / clock, their means inside 99% CIs, the same
    extinction probability, and matching pooled final ecDNA distributions."""
    from scipy import stats as sps
    n = 10000
    o = pkg.SimulationOptions(b0=1.0, b1=1.3, d0=0.1, d1=0.15, cells=3000, runs=n, save_snapshots=False)
    g = ctx.run(o, want=WANT, hist_stride=256)
    ref = ob.run_batch(oracle_opts(o, 0, state=ob.STATE_VECTOR, rng=ob.RNG_RAND), 10 ** 6, n, hist_cap=256)
    ga, ra = g.stop == pkg.STOP_MAX_CELLS, ref.stop == ob.STOP_MAX_CELLS
    pg, pr = 1 - ga.mean(), 1 - ra.mean()
    assert abs(pg - pr) < 2.58 * np.sqrt((pg * (1 - pg) + pr * (1 - pr)) / n)
    rstats = np.array([ob.stats(h) for h in ref.hist[ra]])
    samples = {"mean": (g.mean[ga], rstats[:, 0]), "frequency": (g.frequency[ga], rstats[:, 1]),
               "entropy": (g.entropy[ga], rstats[:, 2]), "clock": (g.time[ga], ref.time[ra])}
    for name, (a, b) in samples.items():
        a, b = a.astype(np.float64), b.astype(np.float64)
        assert sps.ks_2samp(a, b).pvalue > 0.01, name
        se = np.sqrt(a.var() / len(a) + b.var() / len(b))
        assert abs(a.mean() - b.mean()) < 2.58 * se, name
    # pooled final ecDNA distribution: per-class frequencies agree within the between-replicate spread
    hg, hr = g.hist[ga].astype(np.float64), ref.hist[ra].astype(np.float64)
    fg, fr = hg / hg.sum(axis=1, keepdims=True), hr / hr.sum(axis=1, keepdims=True)
    se = np.sqrt(fg.var(axis=0) / len(fg) + fr.var(axis=0) / len(fr)) + 1e-9
    z = np.abs(fg.mean(axis=0) - fr.mean(axis=0)) / se
    busy = (fg.mean(axis=0) + fr.mean(axis=0)) > 1e-3
    assert (z[busy] > 3.5).sum() <= 1 and z[busy].max() < 5.0
    assert np.abs(fg.mean(axis=0).cumsum() - fr.mean(axis=0).cumsum()).max() < 0.008
    # two-sample KS on the final ecDNA distribution with independent samples: one random cell per replicate
    rng = np.random.default_rng(5)

    def one_cell(f):
        u = rng.random(len(f))
        return (f.cumsum(axis=1) < u[:, None]).sum(axis=1)

    assert sps.ks_2samp(one_cell(fg), one_cell(fr)).pvalue > 0.01


def test_full_size_properties(pkg, ctx):
    """BASELINE config 1 at full size (1e5 cells, 100 replicates): size-independent invariants."""
    o = pkg.SimulationOptions(cells=100000, runs=100, save_snapshots=False)
    res = ctx.run(o, want=WANT, digest=True)
    assert np.all(res.stop == pkg.STOP_MAX_CELLS)
    assert np.all(res.n_events == 99999)  # pure birth: one cell per event
    assert np.all(res.nminus + res.nplus == 100000)
    np.testing.assert_array_equal(res.hist.sum(axis=1), 100000)
    np.testing.assert_array_equal(res.hist[:, 0], res.nminus)
    # digest of the final histogram recomputed from the histogram itself
    w = np.array([ob.lib().orc_hist_weight(k) for k in range(512)], dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = (res.hist[:, 1:].astype(np.uint64) * w[None, 1:]).sum(axis=1, dtype=np.uint64)
    np.testing.assert_array_equal(h, res.hash)
    # neutral model: total copies per cell is a martingale with mean 1
    assert abs(res.mean.mean() - 1.0) < 4 * res.mean.std() / 10
    # three replicates checked bit for bit against the oracle
    for i in (0, 57, 99):
        assert_run_equal(res, i, ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=512), 512)


def test_copy_overflow_is_reported_not_fatal(pkg, ctx):
    """k >= 32768 cannot double in u16: the reference panics (proliferation.rs:63-67); here the
    replicate stops with COPY_OVERFLOW and the batch goes on."""
    o = pkg.SimulationOptions(cells=50, runs=4, initial={40000: 1}, segregation="deterministic", save_snapshots=False)
    res = ctx.run(o, want=WANT, state_mode=pkg.STATE_HBM, hist_stride=64)
    assert np.all(res.stop == pkg.STOP_COPY_OVERFLOW)
    ref = ob.run(oracle_opts(o, o.idx_begin), hist_cap=64)
    assert ref.stop_reason == ob.STOP_COPY_OVERFLOW


def test_single_replicate_and_ragged_batches(pkg, ctx):
    """Batch sizes that do not fill a tile group, a warp or a block give the same replicates."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.3, d0=0.1, d1=0.1, cells=2000, runs=67, save_snapshots=False)
    full = ctx.run(o, want=WANT, tile_width=4)
    for n in (1, 3, 9, 33):
        part = ctx.run(o, n_runs=n, want=WANT, tile_width=4)
        for f in ("stop_reason", "nminus", "nplus", "n_events", "kmax", "hist"):
            np.testing.assert_array_equal(getattr(part, f), getattr(full, f)[:n], err_msg=f)
    # a later index range is the same as the tail of a longer one (replicates are keyed by index)
    tail = ctx.run(o, n_runs=7, idx_begin=o.idx_begin + 60, want=WANT)
    np.testing.assert_array_equal(tail.hist, full.hist[60:])


def test_bad_params_are_errors(pkg, ctx):
    with pytest.raises(pkg.EcdnaB200Error):
        ctx.run(pkg.SimulationOptions(b1=-1.0, runs=1, save_snapshots=False))
    with pytest.raises(pkg.EcdnaB200Error):
        ctx.run(pkg.SimulationOptions(runs=1, initial={0: 0}, save_snapshots=False))
    with pytest.raises(pkg.EcdnaB200Error):
        ctx.run(pkg.SimulationOptions(runs=1, save_snapshots=False), tile_width=5)


def _tree(root):
    import os
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = open(p).read()
    return out


def test_cli_chunks_and_devices_give_identical_files(pkg, tmp_path):
    """main.rs:214-225 maps the whole index range over all workers: the CLI runs it on every visible GPU, in
    chunks (host memory O(chunk)); the files do not depend on the chunk size or on the number of GPUs."""
    import os
    import subprocess
    import torch
    pkg.build()
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "host", "ecdna")
    base = ["--b1", "1.3", "--d0", "0.1", "--d1", "0.1", "--cells", "400", "--runs", "23", "--seed", "3", "--subsamples=40",
            "--summaries"]
    trees = []
    variants = [["--devices", "0"], ["--devices", "0", "--chunk", "5"], ["--chunk", "7"]]
    if torch.cuda.device_count() >= 2:
        variants.append(["--devices", "0,1", "--chunk", "8"])
    for i, extra in enumerate(variants):
        out = str(tmp_path / f"o{i}")
        r = subprocess.run([exe] + base + extra + [out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        trees.append(_tree(out))
    assert len(trees[0]) >= 23 * 3
    for t in trees[1:]:
        assert t == trees[0]
    # the literal sosa stop rule (population array [n-,n+,n-,n+] summed: half the size) is one flag away
    out = str(tmp_path / "half")
    r = subprocess.run([exe, "--d0", "0.1", "--cells", "400", "--runs", "2", "--snapshots=", "--bd-count-mode", "sosa-sum", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert all(p.startswith("200cells") for p in _tree(out))


def test_multi_context_matches_single_context(pkg, ctx):
    """ecdna_b200_multi_run: contiguous index blocks per GPU written at their offsets = the one-GPU result."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, d0=0.2, d1=0.1, cells=1500, runs=301, snapshots=[1, 100, 1500])
    want = WANT + ("snap_count", "snap_cells", "snap_time", "snap_hist", "dyn", "dyn_count")
    a = ctx.run(o, want=want, dyn_points=50, dyn_dt=0.2)
    m = pkg.MultiContext()
    b = m.run(o, want=want, dyn_points=50, dyn_dt=0.2)
    assert b.timing.n_finished == 301 and b.timing.total_events == a.timing.total_events
    for f in want:
        x, y = getattr(a, f), getattr(b, f)
        if x.dtype.kind == "f":
            x, y = x.view(np.uint32), y.view(np.uint32)
        np.testing.assert_array_equal(x, y, err_msg=f)
    rates = ctx.abc_draw_priors(seed=1, idx_begin=o.idx_begin, n_runs=301)
    c = ctx.run(o, want=("n_events", "nplus"), rates_per_run=rates)
    d = m.run(o, want=("n_events", "nplus"), rates_per_run=rates)
    np.testing.assert_array_equal(c.n_events, d.n_events)
    m.close()


def test_cli_abc_front_end(pkg, ctx, tmp_path):
    """`ecdna abc --target FILE.json`: abc.csv with every draw in the column order of abc.md:38-55, distances
    equal to the oracle's for the same prior draws, abc_accepted.csv = the rows within the thresholds."""
    import csv
    import json
    import os
    import subprocess
    pkg.build()
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "host", "ecdna")
    o = pkg.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=2000, seed=7, save_snapshots=False)
    target = ob.run(oracle_opts(o, 999), hist_cap=256).hist
    tfile = tmp_path / "target.json"
    tfile.write_text(json.dumps({str(k): int(c) for k, c in enumerate(target) if c}))
    out = str(tmp_path / "abc")
    thr = (0.2, 0.5, 0.5, 0.5)
    r = subprocess.run([exe, "abc", "--target", str(tfile), "-r", "300", "--cells", "2000", "--seed", "7", "--chunk", "128",
                        "--thresholds", ",".join(map(str, thr)), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = list(csv.DictReader(open(os.path.join(out, "abc.csv"))))
    assert list(rows[0].keys()) == pkg.ABC_FIELDS and len(rows) == 300
    assert [int(x["idx"]) for x in rows] == list(range(70, 370))
    rates = np.array([[x["f2"], x["f1"], x["d2"], x["d1"]] for x in rows], dtype=np.float32)  # b0, b1, d0, d1
    np.testing.assert_array_equal(rates, ctx.abc_draw_priors(seed=7, idx_begin=70, n_runs=300))
    ref = ob.abc_batch(oracle_opts(o, 0), 70, 300, rates, target, thr, 0, hist_cap=512)
    got = np.array([[x["ecdna"], x["mean"], x["entropy"]] for x in rows], dtype=np.float32)
    np.testing.assert_array_equal(got.view(np.uint32), ref.distance[:, :3].copy().view(np.uint32))  # (%.9g round-trips f32)
    acc = list(csv.DictReader(open(os.path.join(out, "abc_accepted.csv"))))
    assert {int(x["idx"]) for x in acc} == {70 + i for i in range(300) if ref.accept[i]} and 0 < len(acc) < 300
    assert all(int(x["init_cells"]) == 1 and int(x["init_copies"]) == 1 for x in rows[:5])


def test_lane_cascade_is_invisible_in_the_results(pkg, ctx):
    """1-lane tiles with a 128-bin window (ten warps per SM instead of six): replicates whose copy numbers pass
    127 continue in a 256-bin launch of the same kernel and, beyond 255, in HBM.  Same bits as one 256-bin launch."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.3, cells=20_000, runs=3000, initial={90: 1}, save_snapshots=False)
    a = ctx.run(o, want=WANT, tile_width=1, smem_bins=128)
    b = ctx.run(o, want=WANT, tile_width=1, smem_bins=256)
    assert a.timing.tile_width == 1 and a.timing.smem_bins == 128 and a.timing.kernel_launches == 3
    assert b.timing.smem_bins == 256 and b.timing.kernel_launches == 2
    assert int((a.kmax >= 128).sum()) > 100 and int((a.kmax >= 256).sum()) > 10
    for f in ("stop_reason", "nminus", "nplus", "n_events", "kmax", "hist", "sum_k", "n_div", "n_death"):
        x, y = getattr(a, f), getattr(b, f)
        if f == "stop_reason":
            x, y = x & 0xFF, y & 0xFF
        np.testing.assert_array_equal(x, y, err_msg=f)
    np.testing.assert_array_equal(a.time.view(np.uint32), b.time.view(np.uint32))
    np.testing.assert_array_equal(a.mean.view(np.uint32), b.mean.view(np.uint32))
    for i in (0, int(np.argmax(a.kmax)), 2999):
        assert_run_equal(a, i, ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=512), 512, digest=False)


@pytest.mark.parametrize("name", ["selection", "birth_death", "no_uneven"])
def test_full_launch_build_of_lane_tiles_bit_exact(pkg, ctx, name):
    """A launch of 1-lane tiles with more than one warp per scheduler runs the build of the kernel whose event loop
    is unrolled once (engine.cuh: launch_kernel); pure birth, birth-death and a non-default segregation rule cover
    its three specialisations.  Every replicate against 2-lane tiles (another kernel), some against the oracle;
    snapshots and dynamics included."""
    n = 24_000  # > 4 schedulers x 148 SMs x 32 lanes
    kw = dict(CASES[name], cells=300)
    o = pkg.SimulationOptions(runs=n, snapshots=[1, 40, 200, 300], **kw)
    want = WANT + ("snap_count", "snap_cells", "snap_time", "snap_hist", "dyn", "dyn_count")
    a = ctx.run(o, want=want, tile_width=1, dyn_points=20, dyn_dt=0.25)
    b = ctx.run(o, want=want, tile_width=2, dyn_points=20, dyn_dt=0.25)
    assert a.timing.tile_width == 1 and a.timing.grid_blocks * 2 > 4 * 148 and a.timing.n_finished == n
    for f in want:
        if f in ("hash", "chain"):
            continue
        x, y = getattr(a, f), getattr(b, f)
        if x.dtype.kind == "f":
            x, y = x.view(np.uint32), y.view(np.uint32)
        np.testing.assert_array_equal(x, y, err_msg=f)
    for i in (0, 1, n // 2, n - 1):
        ref = ob.run(oracle_opts(o, o.idx_begin + i, snapshots=o.snapshots, dyn_points=20, dyn_dt=0.25), hist_cap=512)
        assert_run_equal(a, i, ref, 512, digest=False)
        assert_stats_equal(a, i, ref.hist)
        assert int(a.dyn_count[i]) == ref.dyn_count and int(a.snap_count[i]) == ref.n_snap_taken
        np.testing.assert_array_equal(a.dyn[i][:ref.dyn_count].view(np.uint32), ref.dyn[:ref.dyn_count].view(np.uint32))
