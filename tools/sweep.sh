# tuning helper: one line per configuration given as "ENV=VAL ..." strings
for cfg in "X=0"; do
  env $cfg python bench.py --no-e2e --no-cpu --steps 50 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'ms/step=%.4f obs_ms=%.4f frac=%.3f step_ms=%.4f whole=%.3f nominal=%.3f value=%.4e' % (d['ms_per_step'], r['kernel_ms'], r['frac'], r['step_kernel_ms'], r['whole_step']['frac'], r['whole_step']['frac_of_nominal_8TBs'], d['value']))"
done
