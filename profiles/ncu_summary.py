#!/usr/bin/env python
"""Print the headline counters of a `ncu --set full` report (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shfl.sum" if False else "sm__inst_executed_pipe_uniform.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        print("kernel:", d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]:>18s} {units[head.index(k)]}")
        stalls = [(float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") or
                  (k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith(".ratio")) if v not in ("", "n/a")]
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {k:66s} {v:10.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
