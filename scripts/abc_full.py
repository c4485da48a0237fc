"""BASELINE config 4 at full size on one GPU: 1e6 prior draws over (b1, d0, d1), 1e5-cell runs,
distances + accept fused in the kernel epilogue, accepted draws compacted on device."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
import torch

m = _pkg.load()
draws = int(os.environ.get("DRAWS", "1000000"))
ctx = m.Context(0)
dev = torch.device("cuda", 0)
opts = m.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=100_000, runs=draws, save_snapshots=False)
tgt = ctx.run(opts, n_runs=1, idx_begin=260, want=("hist",)).hist[0].astype(np.uint64)
rates = ctx.abc_draw_priors(seed=26, idx_begin=opts.idx_begin, n_runs=draws)
rates_d = torch.from_numpy(rates).to(dev)
tgt_d = torch.from_numpy(tgt.astype(np.int64)).to(dev)
want = ("stop_reason", "n_events", "abc_distance", "abc_accept", "mean", "frequency", "entropy")
rs, t = m.device_results(torch, draws, want, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
torch.cuda.synchronize()
t0 = time.perf_counter()
ctx.run_device(opts, draws, opts.idx_begin, rs, stream=stream, rates_per_run=rates_d, abc_target=tgt_d,
               abc_thresholds=(0.05, 0.1, 0.1, 0.1), slice_events=int(os.environ.get("SLICE", "0"), 0))
acc_idx = torch.empty(draws, dtype=torch.int32, device=dev)
n_acc = ctx.compact_accepted(t["abc_accept"].data_ptr(), draws, acc_idx.data_ptr(), stream=stream)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tm = ctx.timing()
stops = np.bincount(t["stop_reason"].cpu().numpy() & 0xFF, minlength=5)
post = rates_d[acc_idx[:n_acc].long()].cpu().numpy()
print(f"draws={draws} wall={dt:.2f}s kernel={tm.kernel_ms/1e3:.2f}s sims/s={draws/dt:.4g} events={tm.total_events:.4g} "
      f"events/s={tm.total_events/dt:.4g} accepted={n_acc} stops={stops.tolist()} spilled={tm.n_spilled} slice={tm.slice_events} n_slices={tm.n_slices}")
if n_acc:
    print("posterior mean (b1,d0,d1):", post[:, 1:].mean(axis=0), " truth: 1.4 0.2 0.2")
