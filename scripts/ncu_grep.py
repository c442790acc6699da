"""python scripts/ncu_grep.py rep.ncu-rep REGEX  -> metric values (first kernel) whose name fully matches REGEX"""
import csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    for a, b, c in zip(h, u, r):
        if re.fullmatch(sys.argv[2], a):
            print("%-80s %s %s" % (a, c, b))
    print()
