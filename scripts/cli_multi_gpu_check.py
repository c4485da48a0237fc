"""The `ecdna` CLI on every GPU of the box against one GPU: same files (names and contents), wall time and
peak resident memory of the process.  usage: python scripts/cli_multi_gpu_check.py [runs] [cells] [chunk]"""
import hashlib
import os
import resource
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "ecdna-evo_b200", "host", "ecdna")
runs, cells, chunk = (sys.argv[1:] + ["20000", "100000", "4000"])[:3]
args = ["--b1", "1.2", "--d0", "0.1", "--d1", "0.1", "--cells", cells, "--runs", runs, "--snapshots=", "--seed", "5"]


def tree_digest(root):
    h, n = hashlib.sha256(), 0
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            p = os.path.join(d, f)
            h.update(os.path.relpath(p, root).encode())
            h.update(open(p, "rb").read())
            n += 1
    return n, h.hexdigest()


out = {}
for name, extra in (("one GPU, one call", ["--devices", "0", "--chunk", runs]), ("all GPUs, chunks of " + chunk, ["--chunk", chunk])):
    d = f"/tmp/ecdna_cli_{len(out)}"
    subprocess.run(["rm", "-rf", d])
    before = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    t0 = time.time()
    r = subprocess.run([exe] + args + extra + [d], capture_output=True, text=True)
    dt = time.time() - t0
    rss = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    assert r.returncode == 0, r.stderr
    out[name] = tree_digest(d)
    print(f"{name}: {dt:.1f} s, peak RSS of all children so far {rss / 1024:.0f} MiB (before: {before / 1024:.0f}), "
          f"{out[name][0]} files, sha256 {out[name][1][:16]}")
vals = list(out.values())
print("IDENTICAL" if vals[0] == vals[1] else "DIFFERENT")
sys.exit(0 if vals[0] == vals[1] else 1)
