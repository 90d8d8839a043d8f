mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
for wl in c3 c2; do
  timeout 600 python bench.py --workload $wl --no-cpu > gpurun_out/r02b_bench_$wl.json 2> gpurun_out/r02b_bench_$wl.err || tail -5 gpurun_out/r02b_bench_$wl.err
done
GTE_FUSED=0 timeout 600 python bench.py --workload c3 --no-cpu --no-e2e > gpurun_out/r02b_bench_c3_unfused.json 2> gpurun_out/r02b_bench_c3_unfused.err
python - <<'PY'
import json
for wl in ("c3","c2","c3_unfused"):
    try:
        d=json.loads(open(f"gpurun_out/r02b_bench_{wl}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(wl, "FAILED", e); continue
    r=d["roofline"]
    print(wl, "value=%.4e ms=%.5f"%(d["value"],d["ms_per_step"]), "whole=%.3f gather_ms=%.4f step_ms=%.4f launches=%d"%(r["whole_step"]["frac"], r["kernel_ms"], r["step_kernel_ms"], d["gpu_launches"]))
    for k in ("e2e","e2e_gymnasium_dtypes","e2e_other_host_io","e2e_full_obs_to_host"):
        e=d.get(k)
        if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%(1e6*d["config"]["envs_per_gpu"]/e["value"]))
    print("   latency", d.get("latency"))
PY
