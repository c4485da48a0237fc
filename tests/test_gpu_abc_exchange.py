"""The one exchange step of the path (abc.md:57-78) through the C ABI on a GPU: accepted ABC draws
packed into records on the device (ecdna_b200_abc_pack) and all-gathered with the library's own NCCL
communicator (ecdna_b200_comm_* / ecdna_b200_abc_allgather) - a one-rank communicator here, the same code path
bench.py runs on 2-8 ranks."""
import numpy as np
import pytest

import oracle_binding as ob
from test_gpu_parity import oracle_opts

pytestmark = pytest.mark.gpu
BINS = 128


def _abc_batch(pkg, ctx, torch, n, idx_begin, thr):
    dev = torch.device("cuda", 0)
    o = pkg.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=2000, runs=n, save_snapshots=False)
    target = ob.run(oracle_opts(o, 260), hist_cap=BINS).hist
    want = ("stop_reason", "n_events", "nminus", "nplus", "kmax", "abc_distance", "abc_accept", "mean", "frequency",
            "entropy", "hist")
    rs, t = pkg.device_results(torch, n, want, hist_stride=BINS, device=dev)
    rates = torch.empty((n, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        ctx.abc_draw_priors_device(26, idx_begin, n, rates.data_ptr(), stream=stream.cuda_stream)
        ctx.run_device(o, n, idx_begin, rs, stream=stream.cuda_stream, rates_per_run=rates,
                       abc_target=torch.from_numpy(target.astype(np.int64)).to(dev), abc_thresholds=thr, hist_stride=BINS)
    return o, target, rs, t, rates, stream


def test_pack_matches_the_result_columns(pkg, ctx):
    import torch
    n, idx0, thr = 3000, 777, (0.2, 0.5, 0.5, 0.5)
    o, target, rs, t, rates, stream = _abc_batch(pkg, ctx, torch, n, idx0, thr)
    dev = rates.device
    cap = n
    words = pkg.record_words(BINS)
    rec = torch.zeros((cap, words), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx.abc_pack(rs, rates.data_ptr(), (o.b0, o.b1, o.d0, o.d1), idx0, n, BINS, BINS, cap, rec.data_ptr(), cnt.data_ptr(),
                 stream=stream.cuda_stream)
    stream.synchronize()
    acc = np.nonzero(t["abc_accept"].cpu().numpy())[0]
    assert 0 < len(acc) < n and int(cnt.item()) == len(acc)
    d = pkg.decode_records(rec.cpu().numpy().view(np.uint32)[: len(acc)], BINS)
    np.testing.assert_array_equal(d["idx"], idx0 + acc)  # index order
    np.testing.assert_array_equal(d["rates"], rates.cpu().numpy()[acc])
    np.testing.assert_array_equal(d["distance"].view(np.uint32), t["abc_distance"].cpu().numpy()[acc].view(np.uint32))
    for f in ("mean", "frequency", "entropy"):
        np.testing.assert_array_equal(d[f].view(np.uint32), t[f].cpu().numpy()[acc].view(np.uint32), err_msg=f)
    np.testing.assert_array_equal(d["cells"], (t["nminus"] + t["nplus"]).cpu().numpy()[acc])
    np.testing.assert_array_equal(d["kmax"], t["kmax"].cpu().numpy()[acc].astype(np.uint32))
    np.testing.assert_array_equal(d["hist"], t["hist"].cpu().numpy()[acc].astype(np.uint32))
    # ... and the records say what the oracle says about the same draws
    ref = ob.abc_batch(oracle_opts(o, 0), idx0, n, rates.cpu().numpy(), target, thr, 0, hist_cap=BINS)
    np.testing.assert_array_equal(np.nonzero(ref.accept)[0], acc)
    np.testing.assert_array_equal(d["distance"].view(np.uint32), ref.distance[acc].view(np.uint32))
    # a capacity below the count: the count still reports every accepted draw, only `capacity` records are written
    small = torch.zeros((4, words), dtype=torch.int32, device=dev)
    ctx.abc_pack(rs, rates.data_ptr(), (o.b0, o.b1, o.d0, o.d1), idx0, n, BINS, BINS, 4, small.data_ptr(), cnt.data_ptr(),
                 stream=stream.cuda_stream)
    stream.synchronize()
    assert int(cnt.item()) == len(acc)
    np.testing.assert_array_equal(small.cpu().numpy(), rec.cpu().numpy()[:4])
    with pytest.raises(OverflowError):
        pkg.merge_gathered(small.cpu().numpy().view(np.uint32)[None], [len(acc)], 4, BINS)


def test_allgather_through_the_librarys_nccl_communicator(pkg):
    """ecdna_b200_comm_unique_id -> _comm_init -> pack -> _abc_allgather on a communicator of one rank: NCCL is
    found by dlopen, both collectives run on the caller's stream, and the gathered block equals the packed one."""
    import torch
    ctx = pkg.Context(0)
    try:
        with pytest.raises(pkg.EcdnaB200Error):  # no communicator yet
            ctx.abc_allgather(1, 1, BINS, 1, 1, 1)
        ctx.comm_init(pkg.comm_unique_id(), 0, 1)
        n, idx0, thr = 2000, 5000, (0.3, 0.6, 0.6, 0.6)
        o, target, rs, t, rates, stream = _abc_batch(pkg, ctx, torch, n, idx0, thr)
        dev = rates.device
        cap, words = n, pkg.record_words(BINS)
        rec = torch.zeros((cap, words), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        all_rec = torch.full((1, cap, words), -1, dtype=torch.int32, device=dev)
        all_cnt = torch.full((1,), -1, dtype=torch.int32, device=dev)
        ctx.abc_pack(rs, rates.data_ptr(), (o.b0, o.b1, o.d0, o.d1), idx0, n, BINS, BINS, cap, rec.data_ptr(),
                     cnt.data_ptr(), stream=stream.cuda_stream)
        ctx.abc_allgather(rec.data_ptr(), cnt.data_ptr(), BINS, cap, all_rec.data_ptr(), all_cnt.data_ptr(),
                          stream=stream.cuda_stream)
        stream.synchronize()
        n_acc = int(t["abc_accept"].sum().item())
        assert 0 < n_acc <= cap and int(all_cnt.item()) == n_acc == int(cnt.item())
        np.testing.assert_array_equal(all_rec.cpu().numpy()[0], rec.cpu().numpy())
        merged = pkg.merge_gathered(all_rec.cpu().numpy().view(np.uint32), all_cnt.cpu().numpy(), cap, BINS)
        assert merged.shape == (n_acc, words)
        assert np.all(np.diff(pkg.decode_records(merged, BINS)["idx"].astype(np.int64)) > 0)
    finally:
        ctx.close()
