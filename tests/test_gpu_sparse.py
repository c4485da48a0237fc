"""Sparse return of the distributions (ecdna_b200_run_sparse): one descriptor + the occupied window per saved state
(what `save`, reference src/process.rs:31-55, writes is a map over the occupied copy numbers) against the dense
columns of ecdna_b200_run on the same batch, bit for bit, and against the oracle."""
import numpy as np
import pytest

import oracle_binding as ob
from test_gpu_parity import oracle_opts

pytestmark = pytest.mark.gpu

DENSE = ("stop_reason", "nminus", "nplus", "time", "kmax", "hist", "snap_count", "snap_cells", "snap_time", "snap_hist",
         "sub_hist")


def check_against_dense(pkg, dense, res, stride, n_snap, n_sub, one_block=True):
    sp = res.sparse
    n = dense.hist.shape[0]
    # the same simulation
    np.testing.assert_array_equal(dense.stop_reason, res.stop_reason)
    np.testing.assert_array_equal(dense.time.view(np.uint32), res.time.view(np.uint32))
    # every distribution, densified, is the dense column
    np.testing.assert_array_equal(sp.dense(sp.final_dist, stride), dense.hist)
    if n_snap:
        np.testing.assert_array_equal(sp.dense(sp.snap_dist, stride), dense.snap_hist)
    if n_sub:
        np.testing.assert_array_equal(sp.dense(sp.sub_dist, stride), dense.sub_hist)
    # descriptors
    f = sp.final_dist
    np.testing.assert_array_equal(f["cells"], dense.nminus + dense.nplus)
    np.testing.assert_array_equal(f["nminus"], dense.nminus)
    np.testing.assert_array_equal(f["time"].view(np.uint32), dense.time.view(np.uint32))
    assert np.all(f["flags"] & pkg.DIST_TAKEN)
    if n_snap:
        taken = np.arange(n_snap)[None, :] < dense.snap_count[:, None]
        np.testing.assert_array_equal((sp.snap_dist["flags"] & pkg.DIST_TAKEN) != 0, taken)
        np.testing.assert_array_equal(sp.snap_dist["cells"], np.where(taken, dense.snap_cells, 0))
        np.testing.assert_array_equal(sp.snap_dist["time"].view(np.uint32), np.where(taken, dense.snap_time, 0).astype(np.float32).view(np.uint32))
        assert np.all(sp.snap_dist["k_len"][~taken] == 0)
    if n_sub:
        np.testing.assert_array_equal(sp.sub_dist["cells"], dense.sub_hist.sum(axis=2, dtype=np.uint64))
    # the arena is tight: windows back to back in row order (finals, snapshots, samples), first and last bin occupied
    rows = np.concatenate([f.reshape(-1), sp.snap_dist.reshape(-1), sp.sub_dist.reshape(-1)])
    lens = rows["k_len"].astype(np.uint64)
    if one_block:
        np.testing.assert_array_equal(rows["offset"], np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64))
    assert sp.arena_used == int(lens.sum())
    nz = rows[lens > 0]
    by_offset = nz[np.argsort(nz["offset"], kind="stable")]  # (several GPUs: block after block) no gap, no overlap
    np.testing.assert_array_equal(by_offset["offset"], np.concatenate([[0], np.cumsum(by_offset["k_len"].astype(np.uint64))[:-1]]).astype(np.uint64))
    assert np.all(sp.arena[nz["offset"].astype(np.int64)] != 0)
    assert np.all(sp.arena[(nz["offset"] + nz["k_len"] - 1).astype(np.int64)] != 0)
    assert np.all(nz["k_min"] >= 1)
    assert n == f.size


def test_sparse_equals_dense_pure_birth(pkg, ctx):
    o = pkg.SimulationOptions(b0=1.0, b1=1.5, cells=3000, runs=700, seed=5, snapshots=[1, 10, 500, 3000], subsamples=[50, 5000])
    dense = ctx.run(o, want=DENSE)
    res = ctx.run_sparse(o)
    assert res.refetched  # the two-call pattern: size first, then ecdna_b200_sparse_fetch without simulating again
    check_against_dense(pkg, dense, res, 512, 4, 2)
    # far fewer bytes than the dense form: 7 distributions x 512 bins x 4 B per replicate against the occupied bins
    assert res.sparse.arena_used * 4 + 7 * 40 * 700 < dense.hist.nbytes + dense.snap_hist.nbytes + dense.sub_hist.nbytes
    # the final distribution of a few replicates against the oracle itself
    for i in (0, 333, 699):
        ref = ob.run(oracle_opts(o, o.idx_begin + i), hist_cap=512)
        np.testing.assert_array_equal(res.sparse.dense(res.sparse.final_dist[i:i + 1], 512)[0].astype(np.uint64), ref.hist[:512])


def test_sparse_birth_death_with_extinctions(pkg, ctx):
    """Extinct replicates: no window at all, snapshots never reached are marked not taken."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.1, d0=0.6, d1=0.7, cells=800, runs=600, seed=11, snapshots=[1, 5, 100, 800])
    dense = ctx.run(o, want=DENSE[:-1])
    res = ctx.run_sparse(o)
    check_against_dense(pkg, dense, res, 512, 4, 0)
    extinct = (dense.nminus + dense.nplus) == 0
    assert extinct.any() and not extinct.all()
    assert np.all(res.sparse.final_dist["k_len"][extinct] == 0) and np.all(res.sparse.final_dist["cells"][extinct] == 0)
    assert (res.sparse.snap_dist["flags"] == 0).any()


def test_sparse_wide_window(pkg, ctx):
    """{50: 1}: the occupied window sits far from bin 0 (which lives in the descriptor)."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.0, cells=20000, runs=40, seed=2, initial={50: 1}, save_snapshots=False)
    dense = ctx.run(o, want=DENSE[:6], hist_stride=2048)
    res = ctx.run_sparse(o, hist_stride=2048)
    np.testing.assert_array_equal(res.sparse.dense(res.sparse.final_dist, 2048), dense.hist)
    assert res.sparse.final_dist["nminus"].max() > 0 and res.sparse.final_dist["k_min"].min() >= 1


def test_sparse_arena_too_small_then_fetch(pkg, ctx):
    import ctypes as C
    o = pkg.SimulationOptions(b0=1.0, b1=1.3, cells=2000, runs=300, seed=9, snapshots=[100, 2000])
    with pytest.raises(pkg.EcdnaB200Error) as e:
        ctx.run_sparse(o, arena_words=10)
    assert "status 7" in str(e.value) and "needs an arena" in str(e.value)
    # the packed batch is still on the device: fetch it into an arena of the reported size
    sp = pkg.Sparse(300, 2, 0, 0)
    L = pkg.lib()
    assert L.ecdna_b200_sparse_fetch(ctx._h, C.byref(sp.struct)) == pkg.ERR_ARENA
    sp.resize(sp.arena_used)
    assert L.ecdna_b200_sparse_fetch(ctx._h, C.byref(sp.struct)) == 0
    dense = ctx.run(o, want=DENSE[:-1])
    np.testing.assert_array_equal(sp.dense(sp.final_dist, 512), dense.hist)
    np.testing.assert_array_equal(sp.dense(sp.snap_dist, 512), dense.snap_hist)
    # after an ordinary run there is no sparse batch any more
    assert L.ecdna_b200_sparse_fetch(ctx._h, C.byref(sp.struct)) == 1


def test_sparse_many_rows(pkg, ctx):
    """More rows than one pass of the single-block scan over the chunk sums holds (256 chunks of 2048 rows)."""
    o = pkg.SimulationOptions(b0=1.0, b1=1.2, cells=48, runs=60000, seed=4, snapshots=[1, 2, 4, 8, 12, 16, 24, 32, 40, 48])
    dense = ctx.run(o, want=DENSE[:-1], hist_stride=64)
    res = ctx.run_sparse(o, hist_stride=64)
    assert res.sparse.final_dist.size * 11 > 256 * 2048
    check_against_dense(pkg, dense, res, 64, 10, 0)


def test_sparse_multi_context(pkg, ctx):
    """ecdna_b200_multi_run_sparse: every GPU packs its block, the blocks follow each other in the arena."""
    import torch
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]  # (two contexts on one GPU exercise the block bases too)
    m = pkg.MultiContext(devices)
    o = pkg.SimulationOptions(b0=1.0, b1=1.4, d0=0.1, d1=0.1, cells=1500, runs=501, seed=8, snapshots=[1, 100, 1500], subsamples=[64])
    dense = ctx.run(o, want=DENSE)
    res = m.run_sparse(o)
    check_against_dense(pkg, dense, res, 512, 3, 1, one_block=False)
    m.close()
