"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports every symbol that
include/ecdna_b200.h declares, agrees with the ctypes mirror on struct layout, and refuses to run
(rather than falling back) when no B200 is present.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ecdna_b200.h")


@pytest.fixture(scope="module")
def built(pkg):
    pkg.build()
    return pkg


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ecdna_b200_[a-z_0-9]+)\s*\(", src)))


def test_header_functions_are_exported(built):
    names = declared_functions()
    assert len(names) >= 9
    lib = C.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ecdna_b200.h but not exported"
    assert sorted(built.EXPORTED_SYMBOLS) == names
    assert lib.ecdna_b200_abi_version() == built.ABI_VERSION


def test_no_reference_or_oracle_symbols_in_product(built):
    """The product library must not link the oracle: no orc_* symbol, no dependency on it."""
    out = subprocess.run(["nm", "-D", "--defined-only", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out
    ldd = subprocess.run(["ldd", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "torch" not in ldd
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ecdna-evo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_binding" not in txt and "ecdna_oracle" not in txt, f


def test_struct_layout_matches_ctypes(built, tmp_path):
    c = tmp_path / "layout.c"
    c.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "ecdna_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(ecdna_b200_params_t), sizeof(ecdna_b200_results_t), sizeof(ecdna_b200_timing_t), sizeof(ecdna_b200_replay_event_t));
  printf("%zu %zu %zu %zu %zu ", offsetof(ecdna_b200_params_t, max_cells), offsetof(ecdna_b200_params_t, seed), offsetof(ecdna_b200_params_t, init_k), offsetof(ecdna_b200_params_t, abc_thresholds), offsetof(ecdna_b200_params_t, spill_records));
  printf("%zu %zu %zu ", offsetof(ecdna_b200_params_t, slice_events), offsetof(ecdna_b200_params_t, subsample_cells), offsetof(ecdna_b200_results_t, sub_hist));
  printf("%zu %zu %zu ", offsetof(ecdna_b200_results_t, hist), offsetof(ecdna_b200_timing_t, total_events), offsetof(ecdna_b200_timing_t, n_spilled));
  printf("%zu %zu %zu %zu %zu %zu ", sizeof(ecdna_b200_dist_t), offsetof(ecdna_b200_dist_t, offset), offsetof(ecdna_b200_dist_t, time), offsetof(ecdna_b200_dist_t, k_len), offsetof(ecdna_b200_dist_t, k_min), offsetof(ecdna_b200_dist_t, flags));
  printf("%zu %zu %zu\\n", sizeof(ecdna_b200_sparse_t), offsetof(ecdna_b200_sparse_t, arena), offsetof(ecdna_b200_sparse_t, arena_used));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)],
                   check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    P, R, T = built.ParamsT, built.ResultsT, built.TimingT
    want = [C.sizeof(P), C.sizeof(R), C.sizeof(T), built.REPLAY_DTYPE.itemsize,
            P.max_cells.offset, P.seed.offset, P.init_k.offset, P.abc_thresholds.offset, P.spill_records.offset,
            P.slice_events.offset, P.subsample_cells.offset, R.sub_hist.offset,
            R.hist.offset, T.total_events.offset, T.n_spilled.offset]
    D, S = built.DIST_DTYPE, built.SparseT
    want += [D.itemsize] + [D.fields[f][1] for f in ("offset", "time", "k_len", "k_min", "flags")]
    want += [C.sizeof(S), S.arena.offset, S.arena_used.offset]
    assert got == want


def test_header_is_plain_c_with_citations():
    src = open(HEADER).read()
    assert 'extern "C"' in src and "torch" not in src and "std::" not in src
    for cite in ("main.rs:55-211", "main.rs:28-44", "process.rs:31-55", "clap_app.rs:232-238", "proliferation.rs:63-67"):
        assert cite in src


def test_no_gpu_means_error_not_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu tests")
    with pytest.raises(built.EcdnaB200Error) as e:
        built.Context(0)
    assert "status 3" in str(e.value)


def test_launch_plan_without_a_gpu(built):
    """ecdna_b200_plan runs the same planning code as a launch: tile width by batch size, blocks per SM, and
    time slicing only where a batch does not fill a launch evenly."""
    assert built.plan(100) == (32, 1, 100, False)                  # C1: one warp per replicate
    assert built.plan(1000)[0] == 16 and built.plan(2000)[0] == 8  # C5: widest tile within one warp per scheduler
    assert built.plan(4736) == (4, 1, 4736, False)                 # exactly one block of 4-lane tiles per SM
    lanes, w, tiles, sliced = built.plan(10_000)                   # C2: a lane per replicate, 157 blocks of two warps
    assert (lanes, w, tiles, sliced) == (1, 3, 10_048, False)
    assert built.plan(10_000, tile_width=2) == (2, 1, 9472, True)  # 2-lane tiles: one block per SM, sliced
    assert built.plan(9472, tile_width=2) == (2, 1, 9472, False)   # an exact fit needs no slicing
    assert built.plan(10_000, tile_width=4) == (4, 2, 9472, True)
    assert built.plan(10_000, tile_width=4, slice_events=0xFFFFFFFF) == (4, 5, 10_016, False)
    assert built.plan(1_000_000) == (1, 3, 28_416, False)          # C4: many waves, the queue balances them
    assert built.plan(1_000_000, tile_width=2) == (2, 3, 28_416, False)
    assert built.plan(16_384) == (1, 3, 16_384, False)
    with pytest.raises(Exception):
        built.plan(100, tile_width=3)


def test_query_sizes_matches_the_result_buffers(built):
    """ecdna_b200_query_sizes (the two-call pattern of SURVEY 8b): the bytes it reports are exactly the buffers the
    ctypes mirror allocates for the same parameters; no GPU, no context."""
    o = built.SimulationOptions(b1=1.2, d0=0.1, cells=5000, runs=7, subsamples=[10, 20])
    make = built.Context.make_params
    p = make(None, o, 7, dyn_points=30, hist_stride=256, abc_target=__import__("numpy").ones(8, dtype="uint64"))
    sizes = built.query_sizes(p, 7)
    want = tuple(n for n, _, _ in built.RESULT_FIELDS)
    res = built.Results(7, p.n_snapshots, p.dyn_points, 256, want, p.n_subsamples)
    for name, _, _ in built.RESULT_FIELDS:
        assert sizes[name] == getattr(res, name).nbytes, name
    assert sizes["snap_hist"] == 7 * 11 * 256 * 4 and sizes["dyn"] == 7 * 30 * 5 * 4 and sizes["sub_hist"] == 7 * 2 * 256 * 4
    p2 = make(None, built.SimulationOptions(runs=3, save_snapshots=False), 3)
    s2 = built.query_sizes(p2, 3)
    assert s2["snap_hist"] == 0 and s2["dyn"] == 0 and s2["abc_distance"] == 0 and s2["hist"] == 3 * 512 * 4
