mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_baseline_configs.py -m gpu -x -q -k "server or host_paths or strict" 2>&1 | tail -4
timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/r02d_bench_c2.json 2> gpurun_out/r02d_bench_c2.err || tail -20 gpurun_out/r02d_bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02d_bench_c2.json").read().strip().splitlines()[-1])
print("c2 value=%.4e ms=%.5f" % (d["value"], d["ms_per_step"]))
print("latency", {k:v for k,v in d["latency"].items() if k!="note"})
for k in ("e2e","e2e_gymnasium_dtypes","e2e_other_host_io","e2e_server","e2e_full_obs_to_host"):
    e=d.get(k)
    if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%e["us_per_step"])
PY
