"""Per CUDA source line: warp-instructions executed per warp-iteration and stall-sample share.
usage: python scripts/line_costs.py rep.ncu-rep <warp-iterations> [min_instr]"""
import csv, io, subprocess, sys
rep, iters = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
out = []
tot_s = 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0].isdigit() and r[hdr.index("Instructions Executed")].isdigit():
        ie, smp = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
        out.append((int(r[0]), r[1].strip(), ie / iters, smp)); tot_s += smp
acc = 0
for ln, src, n, smp in out:
    if n >= thr:
        acc += n
        print(f"{ln:5d} {n:7.1f} {100*smp/max(tot_s,1):5.1f}%  {src[:110]}")
print("sum of listed:", acc)
