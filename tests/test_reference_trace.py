"""The turnkey check of a REAL reference trace (scripts/check_reference_trace.py),
exercised on a trace the oracle generated itself: it must pass, a corrupted generator stream must fail check A
and a shifted stream must fail checks A and B.  With a trace of the unmodified Rust reference the same script
settles the [RECALL] items R1-R7 of SURVEY 8c."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "check_reference_trace.py")


def test_self_generated_trace_passes_and_corruptions_fail(tmp_path):
    r = subprocess.run([sys.executable, SCRIPT, "--self-test", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PASS") >= 5 and "A must FAIL" in r.stdout


def test_trace_files_round_trip(tmp_path):
    """The file format of the script: trace.json + little-endian trace.u64, checked from the command line."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import check_reference_trace as crt
    import oracle_binding as ob
    meta, log = crt.write_self_trace(ob, str(tmp_path), "t", b1=1.5, d0=0.0, d1=0.0, cells=1500, years=14,
                                     segregation="binomial-no-uneven", idx=777)
    assert json.load(open(tmp_path / "t.json"))["final"] == meta["final"]
    assert np.array_equal(np.fromfile(tmp_path / "t.u64", dtype="<u8"), log)
    ok = subprocess.run([sys.executable, SCRIPT, str(tmp_path / "t.json"), str(tmp_path / "t.u64")], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stdout
    # a wrong final distribution in the JSON (the reference disagrees with the oracle) is reported
    meta["final"]["1"] = meta["final"].get("1", 0) + 1
    json.dump(meta, open(tmp_path / "t.json", "w"))
    bad = subprocess.run([sys.executable, SCRIPT, str(tmp_path / "t.json"), str(tmp_path / "t.u64")], capture_output=True, text=True)
    assert bad.returncode == 1 and "B1" in bad.stdout and "FAIL" in bad.stdout
