# usage: bash tools/r02_ncu_one.sh <tag> <name> <kernel regex> <workload>
TAG=$1; name=$2; k=$3; wl=$4
mkdir -p gpurun_out
python bench.py --workload $wl --no-e2e --no-cpu --no-configs --steps 8 --warmup 3 > gpurun_out/${TAG}_plain_$name.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain_$name.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$k" -s 6 -c 1 -f -o gpurun_out/${TAG}_ncu_$name \
  python bench.py --workload $wl --no-e2e --no-cpu --no-configs --steps 8 --warmup 3 > gpurun_out/${TAG}_ncu_$name.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_ncu_$name.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_${name}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_ncu_${name}*
