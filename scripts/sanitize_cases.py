"""Small launches of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python scripts/sanitize_cases.py
Covers: 1-lane tiles, 2-lane tiles time-sliced (ring + records), 4-lane tiles parking into the HBM launch,
the HBM-resident launch, replay, uniform-level replay, subsamples, dynamics, the ABC epilogue + packing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import _pkg

m = _pkg.load()
ctx = m.Context(0)
W = ("stop_reason", "nminus", "nplus", "time", "n_events", "kmax", "hist")
n = int(os.environ.get("RUNS", "600"))


def case(name, o, **kw):
    r = ctx.run(o, want=kw.pop("want", W), **kw)
    t = r.timing
    print(f"{name}: {t.total_events} events, tile {t.tile_width}, launches {t.kernel_launches}, slices {t.n_slices}, "
          f"spilled {t.n_spilled}, finished {t.n_finished}/{o.runs}")
    assert t.n_finished == o.runs
    return r


bd = dict(b0=1.0, b1=1.2, d0=0.2, d1=0.1, save_snapshots=False)
case("lane tiles", m.SimulationOptions(cells=300, runs=n, **bd), tile_width=1)
case("lane tiles + dynamics + snapshots", m.SimulationOptions(cells=300, runs=200, b1=1.2, d0=0.2, d1=0.1, snapshots=[1, 50, 300]),
     tile_width=1, dyn_points=40, dyn_dt=0.2, want=W + ("dyn", "dyn_count", "snap_count", "snap_hist", "snap_cells", "snap_time"))
case("2-lane sliced", m.SimulationOptions(cells=200, runs=9600, **bd), tile_width=2, slice_events=64)
case("4-lane sliced", m.SimulationOptions(cells=200, runs=5200, **bd), tile_width=4, slice_events=64)
case("parking (128-bin window -> HBM launch)", m.SimulationOptions(cells=1500, runs=64, b1=1.3, initial={100: 1}, save_snapshots=False),
     tile_width=8, smem_bins=128, spill_records=7)
case("HBM-resident", m.SimulationOptions(cells=800, runs=64, b1=1.3, initial={100: 1}, save_snapshots=False), state_mode=m.STATE_HBM)
case("subsamples", m.SimulationOptions(cells=500, runs=32, subsamples=[10, 200, 700], **bd), want=W + ("sub_hist",))
o = m.SimulationOptions(cells=400, runs=256, **bd)
tgt = ctx.run(o, n_runs=1, idx_begin=5, want=("hist",)).hist[0].astype(np.uint64)
rates = ctx.abc_draw_priors(seed=3, idx_begin=o.idx_begin, n_runs=256)
case("ABC epilogue", o, rates_per_run=rates, abc_target=tgt, want=W + ("abc_distance", "abc_accept"))
try:
    import oracle_binding as ob
    oo = ob.make_opts(b0=1.0, b1=1.2, d0=0.2, d1=0.1, state=ob.STATE_VECTOR, rng=ob.RNG_RAND, max_cells=300, run_idx=260)
    r = ob.run(oo, hist_cap=512, trace_cap=4000, u64_cap=200000)
    o1 = m.SimulationOptions(cells=300, runs=1, **bd)
    case("decision replay", o1, replay=r.trace, replay_offsets=np.array([0, len(r.trace)], dtype=np.uint64))
    case("uniform-level replay", o1, replay_u64=r.u64, replay_offsets=np.array([0, len(r.u64)], dtype=np.uint64))
except Exception as e:  # the oracle only generates the streams here
    print("replay cases skipped:", e)
print("all cases ran")
