# instruction count / issue utilisation of kernel regex $1 (two launches)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio \
  --clock-control none -k regex:"${1:-step_kernel}" -s 6 -c 2 --csv --log-file gpurun_out/ncu_inst.csv \
  python bench.py --workload ${WORKLOAD:-c5} --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/ncu_inst.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/ncu_inst.csv") if l.startswith('"')))
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
for r in rows[1:]:
    print(r[ii], r[ki][:30], r[mi], r[vi])
PY
