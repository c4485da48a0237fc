"""Basic-block view of the hot loop of an .ncu-rep (source page): instructions, stall-sample share."""
import csv, io, subprocess, sys
rep = sys.argv[1]; iters = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ia, isrc, iaddr, isamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Address'), hdr.index('# Samples')
hot = [r for r in data if r[ia].isdigit() and int(r[ia]) > 0.02 * iters]
blocks, cur, prev = [], [], None
for r in hot:
    addr = int(r[iaddr], 16)
    op = (r[isrc].split()[1] if r[isrc].startswith('@') else r[isrc].split()[0]).split('.')[0]
    if prev is not None and addr != prev + 16 and cur:
        blocks.append(cur); cur = []
    cur.append(r)
    if op in ('BRA', 'BSYNC', 'EXIT', 'CALL', 'BREAK', 'WARPSYNC', 'BSSY', 'YIELD', 'NOP'):
        blocks.append(cur); cur = []
    prev = addr
if cur: blocks.append(cur)
tot = sum(int(r[isamp]) for r in hot)
print(len(hot), 'hot instr;', sum(int(r[ia]) for r in hot) / iters, 'executed per iteration;', len(blocks), 'blocks')
for b in blocks:
    s = sum(int(r[isamp]) for r in b)
    print(f"{b[0][iaddr][-5:]} n={len(b):3d} time={100*s/tot:5.1f}%  ends: {b[-1][isrc][:60]}")
if len(sys.argv) > 3:
    open(sys.argv[3], 'w').write('\n'.join(f"{r[iaddr][-5:]} {int(r[ia])/iters:5.2f} {r[isamp]:>6s}  {r[isrc]}" for r in hot))
