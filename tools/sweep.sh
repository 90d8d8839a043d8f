# tuning helper: one line per configuration given as "ENV=VAL ..." strings; WORKLOAD selects the bench workload
WORKLOAD=${WORKLOAD:-c5}
for cfg in "X=0" "GTE_TMA_STAGES=4 GTE_TMA_RTILES=2 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=3 GTE_TMA_RTILES=3 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=3 GTE_TMA_RTILES=2 GTE_TMA_GROUP=2" "GTE_TMA_STAGES=6 GTE_TMA_RTILES=2 GTE_TMA_GROUP=2" "GTE_TMA_STAGES=3 GTE_TMA_RTILES=2 GTE_TMA_GROUP=8"; do
  env $cfg python bench.py --workload $WORKLOAD --no-e2e --no-cpu --steps ${STEPS:-100} --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'us/step=%.2f obs_us=%.2f frac=%.3f step_us=%.2f' % (1e3*d['ms_per_step'], 1e3*r['kernel_ms'], r['frac'], 1e3*r['step_kernel_ms']))"
done
