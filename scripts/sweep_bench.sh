#!/bin/bash
# K x lanes sweep of the bench workload: evals/s, kernel ms, validity of the finished run
# usage: scripts/sweep_bench.sh "2500 5000" "8 16 32" [steps]
KS=${1:-"2500 5000"}; LS=${2:-"8 16 32"}; ST=${3:-300}
for K in $KS; do for L in $LS; do
python bench.py --no-cpu-baseline --steps $ST --k $K --lanes $L 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
f=d['full_run'][0]
print('K=$K lanes=$L value %.3e e2e %.3e ms/step %.3f kernel_ms %.3f steps/launch %.0f nsigma %.2f nmcmc %d total_evals %.3e'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['ms_per_launch'],d['roofline']['mcmc_steps_per_launch'],f['nsigma'],f['nmcmc_final'],f['total_evals']))"
done; done
