# round-2 GPU run 1: parity tests on the new host path + baseline numbers + host topology
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | head -40; free -g; cat /sys/devices/system/node/node*/cpulist; nproc; } > gpurun_out/r02_topology.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
for wl in c5 c3 c2; do
  timeout 600 python bench.py --workload $wl --no-cpu > gpurun_out/r02a_bench_$wl.json 2> gpurun_out/r02a_bench_$wl.err || tail -5 gpurun_out/r02a_bench_$wl.err
done
python - <<'PY'
import json
for wl in ("c5","c3","c2"):
    try:
        d=json.loads(open(f"gpurun_out/r02a_bench_{wl}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(wl, "FAILED", e); continue
    r=d["roofline"]
    print(wl, "value=%.4e ms=%.5f"%(d["value"],d["ms_per_step"]), "whole=%.3f gather_ms=%.4f step_ms=%.4f"%(r["whole_step"]["frac"], r["kernel_ms"], r["step_kernel_ms"]))
    for k in ("e2e","e2e_gymnasium_dtypes","e2e_other_host_io","e2e_full_obs_to_host"):
        e=d.get(k)
        if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%(1e6*d["config"]["envs_per_gpu"]/e["value"]))
    print("   latency", d.get("latency"))
PY
