# tuning helper: one line per configuration given as "ENV=VAL ..." strings
for cfg in "GTE_NO_FUSE=1" "GTE_FUSED_STEP_WARPS=1" "GTE_FUSED_STEP_WARPS=2" "GTE_FUSED_STEP_WARPS=3"; do
  env $cfg python bench.py --no-e2e --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'ms/step=%.4f obs_ms=%.4f step_ms=%.4f whole_frac=%.3f value=%.4e' % (d['ms_per_step'], r['kernel_ms'], r['step_kernel_ms'], r['whole_step']['frac'], d['value']))"
done
