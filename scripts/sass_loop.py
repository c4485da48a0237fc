"""Static look at a kernel's hot loop: python scripts/sass_loop.py <lib.so> <mangled-name-substring>
Prints the instruction count of every backward-branch loop of the function and its local-memory ops."""
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if key not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, "instructions:", len(ins), "local ld/st:", sum(1 for _, t in ins if re.search(r"\b(LDL|STL)", t)))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                j = addr_index[tgt]
                body = ins[j:i + 1]
                loops.append((len(body), tgt, a, sum(1 for _, x in body if re.search(r"\b(LDL|STL)", x)),
                              sum(1 for _, x in body if "CALL" in x)))
    for n, tgt, a, loc, calls in sorted(loops, reverse=True)[:6]:
        print(f"  loop 0x{tgt:x}..0x{a:x}: {n} instructions, {loc} local ld/st, {calls} calls")
