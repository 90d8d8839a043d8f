# usage: bash tools/r02_multi.sh N TAG [steps]
N=$1; TAG=$2; STEPS=${3:-100}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/host_io_scaling_probe.py > gpurun_out/${TAG}_ioprobe_n$N.json 2> gpurun_out/${TAG}_ioprobe_n$N.err || tail -5 gpurun_out/${TAG}_ioprobe_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err || tail -20 gpurun_out/${TAG}_bench_n$N.err
python - $N $TAG <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
try:
    p=json.loads(open(f"gpurun_out/{tag}_ioprobe_n{n}.json").read().strip().splitlines()[-1])
    for k,v in p["copies"].items(): print(k, "alone", v["alone_gbs_per_rank"], "together", v["together_gbs_per_rank"], "agg", v["together_aggregate_gbs"], "ms", v["together_ms_max"])
    print(p["lscpu"]); print("numa", p["numa_nodes"], "nproc", p["nproc"], "mem", p["mem_gb"]); print(p["host_bind"][:2]); print(p["topo"])
except Exception as e: print("probe failed", e)
d=json.loads(open(f"gpurun_out/{tag}_bench_n{n}.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value=%.4e ms=%.5f spread=%s" % (d["value"], d["ms_per_step"], d["spread"]))
for k in ("e2e","e2e_gymnasium_dtypes","e2e_pipelined"):
    e=d.get(k)
    if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%e["us_per_step"])
print(d["clocks"])
PY
