# one `ncu --set full` capture of one launch of kernel regex $1 (default obs_tma) -> gpurun_out/ncu_full_$2.{ncu-rep,csv}
K=${1:-obs_tma}; TAG=${2:-gather}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 6 -c 1 -f -o gpurun_out/ncu_full_$TAG \
  python bench.py --workload ${WORKLOAD:-c5} --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/ncu_full_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_full_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_full_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/
