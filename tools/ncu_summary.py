"""Summarise an `ncu --page raw --csv` export (+ optional `--page source --csv`) into a small JSON: the metrics the
judge reads (DRAM bytes, throughputs, occupancy, stall ratios) and the top stall sites of the source page.

    python tools/ncu_summary.py gpurun_out/X_raw.csv [gpurun_out/X_source.csv] > profiles/X.json
"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "local_load_sectors", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, units, data = r, rows[i + 1], rows[i + 2:]
            break
    out = {}
    for r in data:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        out["kernel"] = d["Kernel Name"]
        for k in KEYS:
            if k in d:
                try:
                    out[k] = float(d[k].replace(",", ""))
                except ValueError:
                    out[k] = d[k]
        u = dict(zip(hdr, units))
        out["_units"] = {k: u[k] for k in KEYS if k in u and u[k]}
        break
    if "dram__bytes_read.sum" in out:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += out[k] * scale.get(out["_units"].get(k, "byte"), 1.0)
        out["dram_bytes_total"] = tot
    return out


def source_summary(path, top=14):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    tot_i = sum(int(r[ia]) for r in data)
    tot_s = sum(int(r[isamp]) for r in data) or 1
    ops = {}
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        o = ops.setdefault(op, [0, 0])
        o[0] += int(r[ia]); o[1] += int(r[isamp])
    warps = max(int(r[ia]) for r in data) or 1
    return {"warp_instructions": tot_i, "samples": tot_s, "sass_lines": len(data),
            "instructions_per_warp_of_hottest_line": tot_i / warps,
            "by_opcode_per_warp": {k: round(v[0] / warps, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:16]},
            "stall_samples_by_opcode_pct": {k: round(100 * v[1] / tot_s, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1][1])[:10]},
            "top_stall_sites": [{"samples_pct": round(100 * int(r[isamp]) / tot_s, 1), "sass": r[isrc].strip()[:80]}
                                for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]],
            "local_memory_sass": {"STL": sum(1 for r in data if " STL" in " " + r[isrc] or r[isrc].strip().startswith("STL")),
                                  "LDL": sum(1 for r in data if "LDL" in r[isrc])}}


if __name__ == "__main__":
    out = raw_summary(sys.argv[1])
    if len(sys.argv) > 2:
        out["source_page"] = source_summary(sys.argv[2])
    json.dump(out, sys.stdout, indent=1)
    print()
