mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -6 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err || tail -20 gpurun_out/r02c_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("c5 value=%.4e ms=%.5f spread=%s whole=%.3f gather_ms=%.4f (%.3f) step_ms=%.4f" % (d["value"], d["ms_per_step"], d["spread"], r["whole_step"]["frac"], r["kernel_ms"], r["frac"], r["step_kernel_ms"]))
for k in ("e2e","e2e_gymnasium_dtypes","e2e_other_host_io","e2e_full_obs_to_host"):
    e=d.get(k)
    if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%e["us_per_step"])
c=d["cpu_baseline"]; print("cpu", "%.3e"%c["value"], c["spread"], "one core %.3e"%c["single_core_value"], c["python_reference"] and {k:c["python_reference"].get(k) for k in ("one_core","all_cores")})
for n,l in (d.get("configs") or {}).items():
    if "error" in l: print(n, l); continue
    print(n, "value=%.4e ms=%.5f whole=%.3f static=%s kern_ms=%.4f kfrac=%.3f step_ms=%s launches=%d" % (l["value"], l["ms_per_step"], l["whole_step_frac"], l["whole_step_frac_with_static_window_read"], l["kernel_ms"], l["kernel_frac"], l["step_kernel_ms"], l["launches_per_step"]))
    if "latency" in l: print("    latency", {k:v for k,v in l["latency"].items() if k!="note"})
    for k,e in (l.get("e2e") or {}).items(): print("    ", k, "%.4e"%e["value"], "us/step=%.2f"%e["us_per_step"], e["host_io"], e["action_dtype"])
r=json.loads(open("gpurun_out/r02c_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm value=%.4e spread=%s" % (r["value"], r["spread"]))
PY
