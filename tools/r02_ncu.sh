# ncu --set full captures of the hot kernels (one launch each) + the launch list of the bench command.
# usage: bash tools/r02_ncu.sh <tag>   -> gpurun_out/<tag>_*.{ncu-rep,csv}
TAG=${1:-r02}
mkdir -p gpurun_out
cap() {  # cap <name> <kernel regex> <workload> [skip]
  local name=$1 k=$2 wl=$3 skip=${4:-6}
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -f -o gpurun_out/${TAG}_ncu_$name \
    python bench.py --workload $wl --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/${TAG}_ncu_$name.log 2>&1
  ncu -i gpurun_out/${TAG}_ncu_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_ncu_$name.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_${name}_source.csv 2>/dev/null
}
python bench.py --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/${TAG}_plain_c5.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain_c5.log; exit 1; }
cap c5_gather obs_tma_coop c5
cap c5_step step_kernel c5
cap c3_fused obs_tma_coop c3
cap c4_gather obs_tma_coop c4
cap c4_step step_kernel c4
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${TAG}_launches_c5.csv \
  python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/${TAG}_launches_c5.log 2>&1
ls -la gpurun_out/ | grep ${TAG}_
