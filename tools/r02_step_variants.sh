# step-kernel variants at the C5 shard size: plain vs TMA-staged pipeline (4 and 3 CTAs per SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_fullsize_properties.py tests/test_cuda_edge_cases.py tests/test_cuda_baseline_configs.py -m gpu -x -q -k "not server and not host_paths" 2>&1 | tail -3
for v in "0 0" "1 0" "3 444"; do
  set -- $v
  GTE_STEP_TMA=$1 GTE_STEP_MAX_CTAS=$2 timeout 300 python bench.py --no-e2e --no-cpu --no-configs --steps 100 > gpurun_out/r02_stepvar_$1.json 2> gpurun_out/r02_stepvar.err || tail -3 gpurun_out/r02_stepvar.err
  python - "$1" <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r02_stepvar_{sys.argv[1]}.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("tma=%s: step_ms=%.4f gather_ms=%.4f iter_ms=%.4f value=%.4e spread=%.3f" % (sys.argv[1], r["step_kernel_ms"], r["kernel_ms"], d["ms_per_step"], d["value"], d["spread"]["rel"]))
PY
done
