# usage: bash tools/r02_scale_one.sh N TAG  -> gpurun_out/<TAG>_bench_c5_n<N>.json (+ a one-screen summary)
N=$1; TAG=$2
mkdir -p gpurun_out
export GTE_HOST_SPIN_TIMEOUT_S=15
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_c5_n$N.json 2> gpurun_out/${TAG}_bench_c5_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_c5_n$N.err
python - $N $TAG <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
d=json.loads(open(f"gpurun_out/{tag}_bench_c5_n{n}.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value=%.4e ms=%.5f spread=%s" % (d["value"], d["ms_per_step"], d["spread"]))
for k in ("e2e","e2e_no_relay","e2e_gymnasium_dtypes","e2e_pipelined","e2e_f32_reward"):
    e=d.get(k)
    if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%e["us_per_step"], "d2h", e["d2h_bytes_per_step"])
print((d.get("e2e") or {}).get("relay"))
print(d["clocks"])
PY
