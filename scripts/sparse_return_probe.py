"""Dense against sparse return of the distributions on BASELINE config 2 with the reference's default eleven
snapshots (clap_app.rs:102-134): bytes over PCIe, host memory, time of the whole call (ecdna_b200_timing_t.total_ms:
copies + kernels) and wall clock.  Usage: python scripts/sparse_return_probe.py [cells] [runs]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import _pkg

pkg = _pkg.load()
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
ctx = pkg.Context(0)
o = pkg.SimulationOptions(b0=1.0, b1=1.5, cells=cells, runs=runs)
dense_want = ("stop_reason", "nminus", "nplus", "time", "kmax", "hist", "snap_count", "snap_cells", "snap_time", "snap_hist")
out = {"cells": cells, "runs": runs, "snapshots": len(o.snapshots)}
for stride in (512, 1024):
    for name in ("dense", "sparse"):
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            if name == "dense":
                r = ctx.run(o, want=dense_want, hist_stride=stride)
                host = sum(getattr(r, f).nbytes for f in dense_want)
            else:
                r = ctx.run_sparse(o, want=dense_want[:5], hist_stride=stride)
                sp = r.sparse
                host = sp.final_dist.nbytes + sp.snap_dist.nbytes + sp.arena_used * 4 + sum(getattr(r, f).nbytes for f in dense_want[:5])
            wall = (time.perf_counter() - t0) * 1e3
            t = r.timing
            rec = dict(total_ms=round(t.total_ms, 2), kernel_ms=round(t.kernel_ms, 2), d2h_MB=round(t.d2h_bytes / 1e6, 2),
                       host_MB=round(host / 1e6, 2), wall_ms=round(wall, 1), launches=t.kernel_launches)
            if best is None or rec["wall_ms"] < best["wall_ms"]:
                best = rec
        out[f"{name}_stride{stride}"] = best
    d, s = out[f"dense_stride{stride}"], out[f"sparse_stride{stride}"]
    # same distributions?
    rd = ctx.run(o, n_runs=64, want=dense_want, hist_stride=stride)
    rs = ctx.run_sparse(o, n_runs=64, hist_stride=stride)
    assert np.array_equal(rs.sparse.dense(rs.sparse.snap_dist, stride), rd.snap_hist)
    out[f"ratio_stride{stride}"] = dict(d2h=round(d["d2h_MB"] / s["d2h_MB"], 1), after_kernel_ms=[round(d["total_ms"] - d["kernel_ms"], 2), round(s["total_ms"] - s["kernel_ms"], 2)])
print(json.dumps(out, indent=1))
