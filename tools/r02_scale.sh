# usage: bash tools/r02_scale.sh N TAG [steps]  — the C5 line on N GPUs of one box (torchrun, one rank per GPU)
N=$1; TAG=$2; STEPS=${3:-200}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/${TAG}_bench_c5_n$N.json 2> gpurun_out/${TAG}_bench_c5_n$N.err || tail -20 gpurun_out/${TAG}_bench_c5_n$N.err
python - $N $TAG <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
d=json.loads(open(f"gpurun_out/{tag}_bench_c5_n{n}.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value=%.4e ms=%.5f spread=%.4f" % (d["value"], d["ms_per_step"], d["spread"]["rel"]))
for k in ("e2e","e2e_gymnasium_dtypes","e2e_pipelined"):
    e=d.get(k)
    if e: print("   ",k,"%.4e"%e["value"], e["host_io"], e["action_dtype"], "us/step=%.2f"%e["us_per_step"])
print(d["clocks"])
PY
