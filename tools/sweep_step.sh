# step-kernel grid shape sweep (CTAs cap -> consecutive tiles per CTA)
for cfg in "X=0" "GTE_STEP_MAX_CTAS=2368" "GTE_STEP_MAX_CTAS=1184" "GTE_STEP_MAX_CTAS=592" "GTE_STEP_MAX_CTAS=1776"; do
  env $cfg python bench.py --no-e2e --no-cpu --steps 60 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'us/step=%.2f obs_us=%.2f step_us=%.2f' % (1e3*d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['step_kernel_ms']))"
done
