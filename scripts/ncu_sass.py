"""Per-instruction view of an ncu report's SASS page: executed counts and stall samples.
usage: python scripts/ncu_sass.py rep.ncu-rep [--ranges]  -> prints, per run of instructions with the same executed count,
the number of instructions, the executed count and the samples; then the top stalled instructions."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and v not in ("", "0") and "Not Issued" not in k}
    ins.append(dict(sass=d["Source"].strip(), ex=int(d["Instructions Executed"] or 0), smp=int(d["# Samples"] or 0), st=st))
tot_ex = sum(i["ex"] for i in ins); tot_s = sum(i["smp"] for i in ins)
print("instructions", len(ins), "executed", tot_ex, "samples", tot_s)
# runs of similar executed count
runs = []
for k, i in enumerate(ins):
    if runs and abs(i["ex"] - runs[-1]["ex"]) <= 0.02 * max(runs[-1]["ex"], 1) :
        runs[-1]["n"] += 1; runs[-1]["smp"] += i["smp"]; runs[-1]["sum"] += i["ex"]; runs[-1]["hi"] = k
    else:
        runs.append(dict(lo=k, hi=k, n=1, ex=i["ex"], smp=i["smp"], sum=i["ex"]))
for r in runs:
    if r["sum"] > 0.002 * tot_ex:
        print("[%4d..%4d] n=%3d ex/inst=%9d  %5.2f%% of executed  %5.2f%% of samples   %s" % (
            r["lo"], r["hi"], r["n"], r["ex"], 100 * r["sum"] / tot_ex, 100 * r["smp"] / tot_s, ins[r["lo"]]["sass"][:50]))
print("--- top stalled instructions")
for k, i in sorted(enumerate(ins), key=lambda kv: -kv[1]["smp"])[:40]:
    print("%4d %5.2f%% ex=%9d %-60s %s" % (k, 100 * i["smp"] / tot_s, i["ex"], i["sass"][:60], dict(collections.Counter(i["st"]).most_common(3))))
