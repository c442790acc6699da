"""Static look at a kernel's SASS: backward branches (loops) with their size and opcode mix.
usage: cuobjdump -sass -fun <mangled> lib.so > k.sass; python scripts/sass_loop.py k.sass"""
import re, sys, collections
ins = []
for l in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\S*\s+(?:!?U?P\d,\s*)?`?\(?0x([0-9a-f]+)\)?", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            loops.append((addr[tgt], i))
print("instructions:", len(ins))
for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:6]:
    h = collections.Counter()
    for a, t in ins[lo:hi + 1]:
        op = [o for o in t.split() if not o.startswith("@")][0]
        h[op.split(".")[0] if not op.startswith("IMAD.MOV") else "IMAD.MOV"] += 1
    print("loop [%d..%d] size %d:" % (lo, hi, hi - lo + 1), ", ".join("%s %d" % kv for kv in h.most_common(18)))
