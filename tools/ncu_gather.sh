# dram traffic + duration of the gather and step kernels (one launch each) under ncu; writes gpurun_out/ncu_gather.csv
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__warps_active.avg.per_cycle_active \
  --clock-control none -k regex:"obs_tma|step_kernel" -s 12 -c 4 --csv --log-file gpurun_out/ncu_gather.csv \
  python bench.py --workload ${WORKLOAD:-c5} --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/ncu_gather.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/ncu_gather.csv") if l.startswith('"')))
h = rows[0]
ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
ii = h.index("ID")
for r in rows[1:]:
    print(r[ii], r[ki][:40], r[mi], r[vi])
PY
