# tuning helper: one line per configuration given as "ENV=VAL ..." strings
for cfg in "GTE_TMA_STAGES=3 GTE_TMA_RSTAGES=12 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=4 GTE_TMA_RSTAGES=6 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=4 GTE_TMA_RSTAGES=7 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=3 GTE_TMA_RSTAGES=6 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=5 GTE_TMA_RSTAGES=6 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=5 GTE_TMA_RSTAGES=8 GTE_TMA_GROUP=4" "GTE_TMA_STAGES=6 GTE_TMA_RSTAGES=8 GTE_TMA_GROUP=2" "GTE_TMA_STAGES=6 GTE_TMA_RSTAGES=12 GTE_TMA_GROUP=2" "GTE_TMA_STAGES=8 GTE_TMA_RSTAGES=12 GTE_TMA_GROUP=2"; do
  env $cfg python bench.py --no-e2e --no-cpu --steps 30 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'ms/step=%.4f obs_ms=%.4f frac=%.3f' % (d['ms_per_step'], r['kernel_ms'], r['frac']))"
done
