"""A short single-launch case for ncu: C2's parameters at reduced population size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg

m = _pkg.load()
cells = int(os.environ.get("CELLS", "20000"))
runs = int(os.environ.get("RUNS", "10000"))
tw = int(os.environ.get("TILE", "32"))
b1 = float(os.environ.get("B1", "1.5"))
d0 = float(os.environ.get("D0", "0"))
d1 = float(os.environ.get("D1", "0"))
dyn = int(os.environ.get("DYN", "0"))
init = {int(os.environ.get("K0", "1")): 1}
state = {"auto": m.STATE_AUTO, "hbm": m.STATE_HBM, "smem": m.STATE_SMEM}[os.environ.get("STATE", "auto")]
bins = int(os.environ.get("BINS", "0"))
slice_events = int(os.environ.get("SLICE", "0"), 0)
ctx = m.Context(0)
o = m.SimulationOptions(b0=1.0, b1=b1, d0=d0 or None, d1=d1 or None, cells=cells, runs=runs, save_snapshots=False, initial=init)
for _ in range(int(os.environ.get("REPS", "2"))):
    r = ctx.run(o, want=("stop_reason", "n_events", "kmax", "mean") + (("dyn", "dyn_count") if dyn else ()), dyn_points=dyn, tile_width=tw, state_mode=state, smem_bins=bins,
                slice_events=slice_events)
t = r.timing
print(f"cells={cells} runs={runs} tile={tw} events={t.total_events} kernel_ms={t.kernel_ms:.3f} "
      f"ev/s={t.total_events / t.kernel_ms * 1e3:.4g} blocks/SM={t.blocks_per_sm} grid={t.grid_blocks} kmax={int(r.kmax.max())} "
      f"slice={t.slice_events} n_slices={t.n_slices} idle_spells={t.n_idle_spells}")
