#!/bin/bash
# final 1-GPU sequence of the round: tests, smoke, bench (driver style, default, reference arm), launch list, ncu --set full of the hot kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=r02h
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_t_all.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_s20.log 2>&1; echo "bench s20 rc=$?"
python bench.py > gpurun_out/${TAG}_bench_default.log 2>&1; echo "bench default rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_ncu.log 2>&1
echo "launch list rc=$?"
python scripts/profile_mh.py 10000 4704 80 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mh_chains -s 70 -c 1 -f -o gpurun_out/${TAG}_mh python scripts/profile_mh.py 10000 4704 80 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu full rc=$?"; tail -1 gpurun_out/${TAG}_plain.log
python scripts/logz_validation.py 16 > gpurun_out/${TAG}_logz_validation.log 2>&1; tail -7 gpurun_out/${TAG}_logz_validation.log
python - <<'PY'
import json
for f in ("r02h_bench_s20", "r02h_bench_default"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % f).read().strip().splitlines()[-1])
        r = d["roofline"]; fr = d["full_run"]
        print(f, "value %.3e e2e %.3e ms/step %.3f | kernel %.4f ms %.4f ns/step fp64 %.3f share %.3f | logz_wall %.3f s nsigma %.2f (infl %.2f) failed %d | matched %.3e | cpu %.3e" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], r["ms_per_launch"], r["ns_per_chain_step"], r["frac"], r["kernel_share_of_step"], d["logz_wall_s"],
            fr["nsigma"], fr["nsigma_batch_inflated"], fr["failed_chains"], d["matched"]["value"], d["cpu_baseline"]["value"]))
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -c 600 gpurun_out/${TAG}_bench_ref.log
