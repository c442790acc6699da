"""Compact summary of an ncu --set full report: python scripts/ncu_summary.py rep.ncu-rep [extra-metric-regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
extra = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
WANT = """Kernel Name|Block Size|Grid Size|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|
gpu__time_duration.sum|launch__registers_per_thread|launch__occupancy_limit|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|
sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum$|smsp__issue_active.avg.pct_of_peak_sustained_active|
smsp__thread_inst_executed_per_inst_executed.ratio|sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|uniform|cbu|adu).avg.pct_of_peak_sustained_active|
smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__sass_thread_inst_executed_op_(dfma|dadd|dmul)_pred_on.sum$|
lts__t_bytes.sum$|l1tex__t_bytes.sum$|sm__cycles_elapsed.avg$|smsp__cycles_active.avg$""".replace("\n", "")
want = re.compile(WANT)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
        if want.search(h) or (extra and extra.search(h)):
            if "issue_stalled" in h and float(v or 0) < 0.02:
                continue
            print("%-92s %s %s" % (h, v, u))
    print()
