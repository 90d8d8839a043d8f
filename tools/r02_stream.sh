mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_fullsize_properties.py tests/test_cuda_baseline_configs.py -m gpu -x -q -k "full_size or c4_full or c5_shard or fused" 2>&1 | tail -5
for v in 0 1; do
  GTE_STREAM_FUSED=$v timeout 300 python bench.py --no-e2e --no-cpu --no-configs --steps 100 > gpurun_out/r02_stream_$v.json 2> gpurun_out/r02_stream.err || tail -5 gpurun_out/r02_stream.err
  python - "$v" <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r02_stream_{sys.argv[1]}.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("stream=%s: iter_ms=%.4f value=%.4e whole=%.3f kernel_ms=%.4f step_ms=%s launches=%d spread=%.3f" % (sys.argv[1], d["ms_per_step"], d["value"], r["whole_step"]["frac"], r["kernel_ms"], r["step_kernel_ms"], d["gpu_launches"], d["spread"]["rel"]))
PY
done
GTE_STREAM_FUSED=1 timeout 300 python bench.py --workload c4 --no-e2e --no-cpu --steps 60 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 stream=1 iter_ms=%.4f value=%.4e' % (d['ms_per_step'], d['value']))"
