"""Follow the fall-through path of a kernel's main loop: python scripts/sass_path.py <lib.so> <name-substring> [taken-addr,...]
Starts at the head of the largest backward-branch loop, follows unconditional branches, does not take
conditional ones (except those whose address is listed), stops at the back edge.  Prints an opcode histogram."""
import collections
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
taken = {int(x, 16) for x in sys.argv[3].split(",")} if len(sys.argv) > 3 else set()
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if key not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    idx = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in idx:
            span = a - int(m.group(1), 16)
            if best is None or span > best[0]:
                best = (span, int(m.group(1), 16), a)
    _, head, back = best
    i = idx[head]
    hist = collections.Counter()
    n = 0
    path = []
    if not taken:  # default: the first conditional branch after the loop head's vote is the hot path
        for j in range(idx[head], idx[head] + 12):
            if ins[j][1].startswith("@") and "BRA" in ins[j][1]:
                taken.add(ins[j][0])
                break
    while True:
        a, t = ins[i]
        n += 1
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        hist[op.split(".")[0]] += 1
        path.append((a, t))
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if a == back:
            break
        cond = t.startswith("@") or re.search(r"BRA\S*\s+!?UP\d", t) is not None
        if m and (not cond or a in taken):
            i = idx[int(m.group(1), 16)]
            continue
        i += 1
        if i >= len(ins) or n > 5000:
            break
    print(name, f"loop head 0x{head:x} back edge 0x{back:x}: {n} instructions on the fall-through path")
    print("  ", ", ".join(f"{k}:{v}" for k, v in hist.most_common()))
    for a, t in path:
        if re.search(r"BRA|LDL|STL|CALL|BSSY|BSYNC|WARPSYNC|VOTE", t):
            print(f"   0x{a:x}: {t}")
