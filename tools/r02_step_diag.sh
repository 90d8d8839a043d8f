mkdir -p gpurun_out
timeout 300 python -m pytest "tests/test_cuda_baseline_configs.py::test_fused_single_launch_equals_two_plain_launches" -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head -12
for v in 0 1 2 4 8 15; do
  GTE_STEP_TMA=0 GTE_STEP_DIAG=$v timeout 300 python bench.py --no-e2e --no-cpu --no-configs --steps 60 > gpurun_out/r02_stepdiag_$v.json 2> gpurun_out/r02_stepdiag.err || tail -3 gpurun_out/r02_stepdiag.err
  python - "$v" <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r02_stepdiag_{sys.argv[1]}.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("diag=%s: step_ms=%.4f gather_ms=%.4f iter_ms=%.4f" % (sys.argv[1], r["step_kernel_ms"], r["kernel_ms"], d["ms_per_step"]))
PY
done
GTE_STEP_TMA=1 GTE_STEP_DIAG=15 timeout 300 python bench.py --no-e2e --no-cpu --no-configs --steps 60 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tma=1 diag=15 step_ms=%.4f' % d['roofline']['step_kernel_ms'])"
