import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
m = _pkg.load()
ctx = m.Context(0)
tw = int(os.environ.get("TILE", "0")); dg = os.environ.get("DIGEST", "0") == "1"
cells = int(os.environ.get("CELLS", "200")); runs = int(os.environ.get("RUNS", "3"))
o = m.SimulationOptions(b0=1.0, b1=float(os.environ.get("B1", "1.0")), cells=cells, runs=runs, save_snapshots=False)
r = ctx.run(o, want=("stop_reason", "n_events", "nminus", "nplus", "kmax"), tile_width=tw, digest=dg)
print("tile", tw, "digest", dg, "stop", r.stop.tolist(), "events", r.n_events.tolist(), "launches", r.timing.kernel_launches, flush=True)
