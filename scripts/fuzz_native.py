"""One-off differential fuzz of the native stream: random parameter sets (rates incl. zeros, tiny and huge values,
all segregation rules, sparse initial distributions, size or time stops, snapshots, dynamics) on every tile width
against the CPU oracle, bit for bit.  usage: python scripts/fuzz_native.py [cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg
import oracle_binding as ob

m = _pkg.load()
ctx = m.Context(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
rules = ["binomial", "deterministic", "binomial-no-uneven", "binomial-no-nminus"]
WANT = ("stop_reason", "nminus", "nplus", "time", "n_events", "kmax", "hist", "mean", "frequency", "entropy", "variance",
        "snap_count", "snap_cells", "snap_time", "snap_hist", "dyn", "dyn_count")
bad = 0
for case in range(cases):
    rate = lambda: float(rng.choice([0.0, 1e-9, 3e-5, 0.3, 0.5, 1.0, 1.7, 3.0, 40.0, 2e8, 7e9]))
    b0, b1, d0, d1 = rate(), rate(), float(rng.choice([0.0, 0.0, rate()])), float(rng.choice([0.0, 0.0, rate()]))
    ks = rng.choice(np.arange(1, 90), size=3, replace=False)
    init = {int(k): int(c) for k, c in zip(ks, rng.integers(1, 6, size=3))}
    if rng.random() < 0.5:
        init[0] = int(rng.integers(1, 8))
    kw = dict(b0=b0, b1=b1, d0=d0, d1=d1, segregation=str(rng.choice(rules)), initial=init, seed=int(rng.integers(1, 1000)))
    if rng.random() < 0.3:
        kw["years"] = int(rng.integers(1, 6))
    else:
        kw["cells"] = int(rng.integers(30, 1500))
    snaps = sorted(set(int(x) for x in rng.integers(1, 1500, size=4))) if rng.random() < 0.5 else None
    runs = int(rng.choice([3, 40, 700]))
    o = m.SimulationOptions(runs=runs, snapshots=snaps, save_snapshots=snaps is not None, **kw)
    if "years" in kw:  # (a time stop alone lets a fast-growing population run to 1e9 cells: cap it)
        o.max_cells = int(rng.integers(30, 1500))
    dyn = dict(dyn_points=int(rng.integers(5, 60)), dyn_dt=float(rng.choice([0.05, 0.3, 1.0]))) if rng.random() < 0.5 else {}
    for tw in (1, 2, 4, 16, 32, 0):
        try:
            res = ctx.run(o, want=WANT, tile_width=tw, **dyn)
        except Exception as e:
            print("case", case, kw, "tile", tw, "ERROR", e); bad += 1; continue
        for i in sorted(set([0, runs // 2, runs - 1])):
            oo = ob.make_opts(b0=o.b0, b1=o.b1, d0=o.d0, d1=o.d1, segregation=o.segregation, state=ob.STATE_HIST, rng=ob.RNG_PHILOX,
                              max_cells=o.max_cells, max_time=float(o.years), seed=o.seed, run_idx=o.idx_begin + i, initial=o.distribution,
                              snapshots=o.snapshots or None, **dyn)
            ref = ob.run(oo, hist_cap=512)
            ok = (int(res.stop[i]) == ref.stop_reason and int(res.n_events[i]) == ref.n_events and int(res.nminus[i]) == ref.nminus
                  and int(res.nplus[i]) == ref.nplus and np.float32(res.time[i]).view(np.uint32) == np.float32(ref.time).view(np.uint32)
                  and int(res.kmax[i]) == ref.kmax and np.array_equal(res.hist[i].astype(np.uint64), ref.hist))
            st = np.array(ob.stats(ref.hist), dtype=np.float32).view(np.uint32)
            got = np.array([res.mean[i], res.frequency[i], res.entropy[i], res.variance[i]], dtype=np.float32).view(np.uint32)
            ok = ok and np.array_equal(st, got)
            if snaps is not None:
                n = ref.n_snap_taken
                ok = ok and int(res.snap_count[i]) == n and np.array_equal(res.snap_hist[i][:n].astype(np.uint64), ref.snap_hist[:n])
            if dyn:
                n = ref.dyn_count
                ok = ok and int(res.dyn_count[i]) == n and np.array_equal(res.dyn[i][:n].view(np.uint32), ref.dyn[:n].view(np.uint32))
            if not ok:
                print("MISMATCH case", case, kw, "runs", runs, "tile", tw, "replicate", i, "snaps", snaps, "dyn", dyn,
                      "| gpu", int(res.stop[i]), int(res.n_events[i]), int(res.nminus[i]), int(res.nplus[i]), "| oracle", ref.stop_reason, ref.n_events, ref.nminus, ref.nplus)
                bad += 1
    if case % 10 == 9:
        print(f"... {case + 1} cases, {bad} mismatches so far", flush=True)
print(f"{cases} cases x 6 tile widths: {bad} mismatches")
sys.exit(1 if bad else 0)
