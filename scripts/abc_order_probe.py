"""Kernel time of a C4-shaped batch with and without the longest-first order (ECDNA_B200_KEEP_ORDER)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
m = _pkg.load()
ctx = m.Context(0)
dev = torch.device("cuda", 0)
for draws in [int(x) for x in os.environ.get("DRAWS", "125000,1000000").split(",")]:
    opts = m.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=100_000, runs=draws, save_snapshots=False)
    rates = torch.from_numpy(ctx.abc_draw_priors(seed=26, idx_begin=opts.idx_begin, n_runs=draws)).to(dev)
    rs, t = m.device_results(torch, draws, ("stop_reason", "n_events"), device=dev)
    for keep in (0, 1, 0, 1):
        p = ctx.make_params(opts, draws, rates_per_run=rates, smem_bins=int(os.environ.get("BINS", "0")),
                            tile_width=int(os.environ.get("TILE", "0")))
        p.flags |= 2 if keep else 0
        import ctypes as C
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx._check(m.lib().ecdna_b200_run_device(ctx._h, C.byref(p), opts.idx_begin, draws, C.byref(rs), None))
        t1 = time.perf_counter()
        tm = ctx.timing()
        t2 = time.perf_counter()
        print(f"  host: enqueue {1e3*(t1-t0):.1f} ms, until done {1e3*(t2-t0):.1f} ms")
        ev = t["n_events"].cpu().numpy()
        print(f"draws={draws} keep_order={keep} kernel_ms={tm.kernel_ms:.1f} events={tm.total_events:.4g} ev/s={tm.total_events/tm.kernel_ms*1e3:.4g} "
              f"max_events={int(ev.max())} p99={int(np.percentile(ev,99))} median={int(np.median(ev))}")
