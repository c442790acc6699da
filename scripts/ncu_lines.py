"""Stall samples per CUDA source line from an ncu report: python scripts/ncu_lines.py rep.ncu-rep [top]"""
import csv, subprocess, sys, collections
def num(v):
    try: return int(v)
    except ValueError: return 0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fpath = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    iN = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
    key = (fpath, r[0])
    a = agg.setdefault(key, dict(src=r[1].strip(), n=0, i=0, st=collections.Counter()))
    if r[1].strip(): a["src"] = r[1].strip()
    a["n"] += num(r[iN]); a["i"] += num(r[iI])
    for k, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[k] not in ("", "0"):
            a["st"][h[6:]] += num(r[k])
    if len(agg) > 1 and fpath and False: break
ts = sum(a["n"] for a in agg.values()); ti = sum(a["i"] for a in agg.values())
print("samples", ts, "instructions", ti)
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    print("%5.2f%% smp %5.2f%% ins %s:%s %-72s %s" % (100 * a["n"] / ts, 100 * a["i"] / ti, f, l, a["src"][:72], dict(a["st"].most_common(3))))
