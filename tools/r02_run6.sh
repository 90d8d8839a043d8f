mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -8 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/r02e_bench_c2.json 2> gpurun_out/r02e_bench_c2.err || tail -20 gpurun_out/r02e_bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02e_bench_c2.json").read().strip().splitlines()[-1])
print("c2 value=%.4e ms=%.5f repeats=%s" % (d["value"], d["ms_per_step"], d["repeats"]))
for k in ("e2e","e2e_gymnasium_dtypes","e2e_other_host_io","e2e_server","e2e_full_obs_to_host"):
    e=d.get(k)
    if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%e["us_per_step"])
PY
