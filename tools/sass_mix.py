"""Instruction mix of the hottest loop of a kernel from `cuobjdump -sass` text (no GPU needed).

    cuobjdump -sass -fun <mangled> file.o > k.sass ; python tools/sass_mix.py k.sass [lo hi]

Finds the innermost backward branch spanning the most instructions that contain the given marker
mnemonic (default VIMNMX3 / FMNMX3), prints the mnemonic histogram inside that loop."""
import re, sys, collections
rx = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)([.\w]*)\s*(.*?);")
ins = []
for line in open(sys.argv[1]):
    m = rx.search(line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), m.group(3), m.group(4)))
addr = {a: i for i, (a, _, _, _) in enumerate(ins)}
loops = []
for i, (a, mn, mod, ops) in enumerate(ins):
    if mn == "BRA":
        t = re.search(r"0x([0-9a-f]+)", ops)
        if t and int(t.group(1), 16) in addr and int(t.group(1), 16) <= a:
            loops.append((addr[int(t.group(1), 16)], i))
marker = sys.argv[2] if len(sys.argv) > 2 else "VIMNMX3|FMNMX3"
best = None
cands = []
for lo, hi in loops:
    n = sum(1 for k in range(lo, hi + 1) if re.match(marker, ins[k][1]))
    if n:
        cands.append((lo, hi, n))
# the hot loop: most marker instructions per loop among loops that contain no other marker loop
for lo, hi, n in cands:
    inner = [c for c in cands if c != (lo, hi, n) and c[0] >= lo and c[1] <= hi]
    print("loop %d..%d: %d instr, %d marker%s" % (lo, hi, hi - lo + 1, n, "" if not inner else " (has inner loops)"))
    if not inner and (best is None or n > best[2]):
        best = (lo, hi, n)
if best is None:
    print("no loop with marker"); sys.exit(1)
lo, hi, n = best
print("loop: instr %d..%d (%d instructions), %d marker instr" % (lo, hi, hi - lo + 1, n))
h = collections.Counter(ins[k][1] + (ins[k][2] if ins[k][1] in ("VIADD", "VIADDMNMX", "VIMNMX3", "VIMNMX", "LDS", "IMAD", "PRMT", "SHFL") else "") for k in range(lo, hi + 1))
for k, v in h.most_common():
    print("%6d  %s" % (v, k))
