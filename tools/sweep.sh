# tuning sweep on the C5 shard: TMA gather pipeline shape (stages x warps per CTA x envs per group)
for cfg in "3 2 4" "3 1 8" "3 4 2" "4 2 2" "3 1 4" "4 1 4" "3 2 2" "3 8 1" "6 2 2" "3 1 2" "6 1 4" "4 4 2" "3 2 8"; do
  set -- $cfg
  GTE_TMA_STAGES=$1 GTE_TMA_WARPS=$2 GTE_TMA_GROUP=$3 python bench.py --no-e2e --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('stages=$1 warps=$2 group=$3', 'ms/step=%.4f obs_ms=%.4f frac=%.3f value=%.4e' % (d['ms_per_step'], r['kernel_ms'], r['frac'], d['value']))"
done
