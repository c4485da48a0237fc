"""The host-side mirror of the reference's option handling and file layout (no GPU needed)."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_years_and_defaults(pkg):
    # clap_app.rs:148-151: default 1000 cells, years = (log2(cells) + 4) as u64
    o = pkg.SimulationOptions()
    assert (o.max_cells, o.years, o.runs, o.seed, o.idx_begin) == (1000, 13, 12, 26, 260)
    assert o.distribution == {1: 1} and not o.birth_death and o.segregation == pkg.SEG_BINOMIAL
    assert pkg.SimulationOptions(cells=100000).years == 20
    assert pkg.SimulationOptions(cells=1000000).years == 23
    assert pkg.SimulationOptions(cells=10000000).years == 27
    y = pkg.SimulationOptions(years=7)  # clap_app.rs:142-147
    assert (y.max_cells, y.years) == (1_000_000_000, 7)


def test_birth_death_detection(pkg):
    # clap_app.rs:165-174, 194-200
    assert not pkg.SimulationOptions(d0=0.0, d1=0.0).birth_death
    assert pkg.SimulationOptions(d0=0.1).birth_death and pkg.SimulationOptions(d1=0.3).birth_death


def test_default_snapshots(pkg):
    # clap_app.rs:121-134
    assert pkg.build_snapshots_from_cells(11, 1000) == [1, 101, 201, 301, 401, 501, 601, 701, 801, 901, 1000]
    assert pkg.build_snapshots(1000, [500, 3, 40]) == [3, 40, 500]
    assert pkg.build_snapshots_from_cells(11, 100000)[1] == 10001


def test_filenames(pkg):
    # lib.rs:27-45: Rust f32 Display with '.' -> 'dot'
    assert pkg.create_filename_pure_birth([1.0, 1.5], 260) == "1b0_1dot5b1_0d0_0d1_260idx"
    assert pkg.create_filename_birth_death([1.0, 1.2, 0.3, 0.3], 261) == "1b0_1dot2b1_0dot3d0_0dot3d1_261idx"
    assert pkg.create_filename_birth_death([1.0, 1.05, 0.0, 0.1], 7) == "1b0_1dot05b1_0d0_0dot1d1_7idx"
    o = pkg.SimulationOptions(b1=1.5)
    assert o.filename(260) == "1b0_1dot5b1_0d0_0d1_260idx"


def test_save_layout(pkg, tmp_path):
    # process.rs:39-45: <dir>/<cells>cells/ecdna/<t>years/<filename>.json ; dynamics.md:7-8 JSON shape
    hist = np.zeros(32, dtype=np.uint32)
    hist[[0, 1, 10, 20]] = [2, 2, 1, 1]
    p = pkg.save(str(tmp_path), "1b0_1b1_0d0_0d1_260idx", 3.14159, hist)
    assert p == os.path.join(str(tmp_path), "6cells", "ecdna", "3dot1years", "1b0_1b1_0d0_0d1_260idx.json")
    assert json.load(open(p)) == {"0": 2, "1": 2, "10": 1, "20": 1}


def test_abc_csv_schema(pkg, tmp_path):
    # abc.md:38-55: one row per draw, all draws kept; f1/d1 = cells with ecDNA, f2/d2 = without
    import csv
    o = pkg.SimulationOptions(b0=1.0, b1=1.4, d0=0.2, d1=0.2, cells=1000, initial={2: 3, 0: 1})
    rates = np.array([[1.0, 1.3, 0.1, 0.2], [1.0, 1.7, 0.3, 0.4]], dtype=np.float32)
    dist = np.array([[0.05, 0.1, 0.2, 0.3], [0.5, 0.6, 0.7, 0.8]], dtype=np.float32)
    p = str(tmp_path / "abc.csv")
    assert pkg.write_abc_csv(p, o, 260, rates, dist, [10, 0], [990, 0]) == 2
    rows = list(csv.DictReader(open(p)))
    assert list(rows[0].keys()) == pkg.ABC_FIELDS
    assert rows[0]["idx"] == "260" and rows[1]["idx"] == "261" and rows[0]["seed"] == "26"
    assert abs(float(rows[0]["f1"]) - 1.3) < 1e-6 and abs(float(rows[0]["f2"]) - 1.0) < 1e-6
    assert abs(float(rows[0]["d1"]) - 0.2) < 1e-6 and abs(float(rows[0]["d2"]) - 0.1) < 1e-6
    assert rows[0]["tumour_cells"] == "1000" and rows[1]["tumour_cells"] == "0"
    assert float(rows[0]["init_mean"]) == 1.5 and rows[0]["init_cells"] == "4" and rows[0]["init_copies"] == "6"
    assert abs(float(rows[1]["ecdna"]) - 0.5) < 1e-6


def test_cli_without_a_gpu_fails_loudly_and_parses_like_clap(tmp_path):
    """The `ecdna` binary has no CPU path: without a B200 it exits 101 (the reference's panic code) with a message
    and writes nothing; argument errors are caught before any device is touched and exit 2 like clap."""
    import subprocess
    import torch
    import _pkg
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the CLI is covered by the gpu tests")
    m = _pkg.load()
    m.build()
    cli = m.CLI_PATH if hasattr(m, "CLI_PATH") else os.path.join(os.path.dirname(m.__file__), "host", "ecdna")
    out = tmp_path / "out"
    r = subprocess.run([cli, "--runs", "2", "--cells", "100", str(out)], capture_output=True, text=True)
    assert r.returncode == 101 and "no CPU path" in r.stderr + r.stdout
    assert not out.exists() or not any(out.rglob("*.json"))
    for bad in (["--segregation", "nonsense", str(out)], [], ["abc", str(out)]):
        r = subprocess.run([cli] + bad, capture_output=True, text=True)
        assert r.returncode == 2 and "error:" in r.stderr
    r = subprocess.run([cli, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and "0.26.0" in r.stdout
