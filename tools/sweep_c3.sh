# small-N shape sweep of the gather (C3)
for cfg in "X=0" "GTE_TMA_STAGES=3 GTE_TMA_RTILES=2 GTE_TMA_GROUP=2" "GTE_TMA_STAGES=3 GTE_TMA_RTILES=2 GTE_TMA_GROUP=1" "GTE_STEP_MIN_CTAS=3"; do
  env $cfg python bench.py --workload c3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'us/step=%.2f obs_us=%.2f step_us=%.2f whole=%.3f' % (1e3*d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['step_kernel_ms'], r['whole_step']['frac']))"
done
