# tuning sweep on the C5 shard: L2 cache-hint combinations of the TMA gather (bit0 obs stores evict_first,
# bit1 window-table loads evict_last, bit2 ring loads evict_first)
for h in 0 1 2 3 4 5 7; do
  GTE_CHUNKS=1 GTE_TMA_HINTS=$h python bench.py --no-e2e --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('hints=$h', 'ms/step=%.4f obs_ms=%.4f step_ms=%.4f whole_frac=%.3f value=%.4e' % (d['ms_per_step'], r['kernel_ms'], r['step_kernel_ms'], r['whole_step']['frac'], d['value']))"
done
