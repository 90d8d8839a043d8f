# tuning helper: static split vs claimed tiles in the TMA gather
for WORKLOAD in c5 c4; do
for cfg in "GTE_TMA_DYN=0" "GTE_TMA_DYN=1"; do
  env $cfg python bench.py --workload $WORKLOAD --no-e2e --no-cpu --no-configs --steps ${STEPS:-100} --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$WORKLOAD $cfg', 'us/step=%.2f obs_us=%.2f frac=%.3f step_us=%.2f clocks=%s' % (1e3*d['ms_per_step'], 1e3*r['kernel_ms'], r['frac'], 1e3*r['step_kernel_ms'], d['clocks']['sm_mhz']))"
done; done
