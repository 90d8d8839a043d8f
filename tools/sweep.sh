# tuning sweep of the TMA gather pipeline shape (stages x warps) on the C5 shard; prints one line per config
for w in 8 4; do for s in 2 3; do
  GTE_TMA_STAGES=$s GTE_TMA_WARPS=$w python bench.py --no-e2e --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('stages=$s warps=$w', 'ms/step=%.4f obs_ms=%.4f frac=%.3f value=%.4e' % (d['ms_per_step'], r['kernel_ms'], r['frac'], d['value']))"
done; done
