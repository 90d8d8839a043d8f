# step-kernel variants at the C5 shard size: software pipeline on/off x CTAs per SM
mkdir -p gpurun_out
GTE_STEP_PIPE=1 timeout 600 python -m pytest tests/test_cuda_fullsize_properties.py tests/test_cuda_edge_cases.py -m gpu -x -q 2>&1 | tail -3
for v in "0 4" "1 4" "1 3" "1 2"; do
  set -- $v
  GTE_STEP_PIPE=$1 GTE_STEP_CTAS_PER_SM=$2 timeout 300 python bench.py --no-e2e --no-cpu --steps 100 > gpurun_out/r02_stepvar_$1_$2.json 2> gpurun_out/r02_stepvar.err || tail -3 gpurun_out/r02_stepvar.err
  python - "$1" "$2" <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r02_stepvar_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("pipe=%s ctas/sm=%s: step_ms=%.4f gather_ms=%.4f iter_ms=%.4f value=%.4e" % (sys.argv[1], sys.argv[2], r["step_kernel_ms"], r["kernel_ms"], d["ms_per_step"], d["value"]))
PY
done
