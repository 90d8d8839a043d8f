# refresh profiles/<round>_bench_*.json on one GPU (the multi-GPU lines are produced by torchrun separately)
R=${ROUND:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/${R}_bench_c5_n1.json 2> gpurun_out/bench_c5.err
python bench.py --workload c3 > gpurun_out/${R}_bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c2 > gpurun_out/${R}_bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --workload c2 --cuda-graph --no-cpu > gpurun_out/${R}_bench_c2_graph.json 2> gpurun_out/bench_c2g.err
python bench.py --workload c4 --no-cpu > gpurun_out/${R}_bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --impl reference > gpurun_out/${R}_bench_ref.json 2> gpurun_out/bench_ref.err
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value=%.4e" % d["value"], "ms=%.4f" % d["ms_per_step"], "e2e=%.4e" % d["e2e"]["value"],
      "frac=%s whole=%s kern_ms=%s step_ms=%s" % (r.get("frac"), (r.get("whole_step") or {}).get("frac"), r.get("kernel_ms"), r.get("step_kernel_ms")),
      "cpu=%s" % (d.get("cpu_baseline") or {}).get("value"), "clocks=%s" % d.get("clocks"))
PY
done
