#!/usr/bin/env python
"""Turn the gpurun_out/ files of `tools/profile_round.sh <tag>` into the committed profiles/<tag>_* summaries."""
import collections
import csv
import io
import json
import os
import shutil
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def last_json(path):
    return [l for l in open(path) if l.startswith("{")][-1]


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main(tag):
    os.chdir(R)
    open(f"profiles/{tag}_bench.json", "w").write(last_json(f"gpurun_out/bench_{tag}.json"))
    open(f"profiles/{tag}_bench_reference.json", "w").write(last_json(f"gpurun_out/bench_ref_{tag}.json"))
    shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches.csv")
    shutil.copy(f"gpurun_out/ncu_{tag}_raw_TomW.csv", f"profiles/{tag}_ncu_raw_wave_TomW.csv")
    d = json.loads(last_json(f"gpurun_out/bench_{tag}.json"))
    r = json.loads(last_json(f"gpurun_out/bench_ref_{tag}.json"))
    print("ours", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "ref", r["value"], "e2e ratio", d["e2e"]["value"] / r["value"])
    traffic, out, mix = {}, io.StringIO(), io.StringIO()
    out.write("ncu --set full --clock-control none --import-source on, one launch (the 3rd, -s 2 -c 1) of each back-end kernel, bench.py --steps 1 --warmup 1 (C2), B200\n")
    for k in ["TomW", "HatW", "SnareW", "KickW"]:
        rows = list(csv.reader(open(f"gpurun_out/ncu_{tag}_raw_{k}.csv")))
        head, units = rows[0], rows[1]
        dd, u = dict(zip(head, rows[2])), dict(zip(head, units))
        out.write(f"\n== wave_kernel<{k}>\n")
        for key in KEYS:
            if key in dd:
                out.write(f"  {key:66s} {dd[key]:>18s} {u[key]}\n")
        st = [(float(v), kk) for kk, v in dd.items() if kk.startswith("smsp__average_warps_issue_stalled") and kk.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
        for v, kk in sorted(st, reverse=True)[:6]:
            out.write(f"  stall {kk[34:-23]:40s} {v:.2f}\n")
        rd, wr = to_bytes(dd["dram__bytes_read.sum"], u["dram__bytes_read.sum"]), to_bytes(dd["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        traffic[f"wave_kernel<{k}>"] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                        "source": f"profiles/{tag}_ncu_summary.txt (ncu --set full, launch 3 of bench.py --steps 1 --warmup 1)"}
        head = None
        m, sm, tot = collections.Counter(), collections.Counter(), 0
        for row in csv.reader(open(f"gpurun_out/ncu_{tag}_source_{k}.csv")):
            if not row:
                continue
            if row[0] == "Address":
                head = row
                continue
            if head is None or len(row) < 7:
                continue
            x = dict(zip(head, row))
            try:
                n, s_ = int(x["Instructions Executed"]), int(x["# Samples"])
            except ValueError:
                continue
            op = x["Source"].split()[0]
            if op.startswith("@"):
                op = x["Source"].split()[1]
            op = op.split(".")[0]
            m[op] += n; sm[op] += s_; tot += n
        ts = sum(sm.values()) or 1
        mix.write(f"== wave_kernel<{k}>: {tot} warp instructions; opcode mix (share of instructions / share of stall samples)\n")
        for op, n in m.most_common(12):
            mix.write(f"   {op:10s} {100 * n / tot:5.1f}%  {100 * sm[op] / ts:5.1f}%\n")
    for k in ["chain_fast_kernel", "mix_kernel", "bass_wave_kernel", "gran_wave_kernel", "coop_kernel"]:      # engine-level kernels and the cooperative back end
        path = f"gpurun_out/ncu_{tag}_raw_{k}.csv"
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        dd, u = dict(zip(head, rows[2])), dict(zip(head, units))
        out.write(f"\n== {k} ({dd.get('Kernel Name', '')[:70]})\n")
        for key in KEYS:
            if key in dd:
                out.write(f"  {key:66s} {dd[key]:>18s} {u[key]}\n")
        st = [(float(v), kk) for kk, v in dd.items() if kk.startswith("smsp__average_warps_issue_stalled") and kk.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
        for v, kk in sorted(st, reverse=True)[:6]:
            out.write(f"  stall {kk[34:-23]:40s} {v:.2f}\n")
        rd, wr = to_bytes(dd["dram__bytes_read.sum"], u["dram__bytes_read.sum"]), to_bytes(dd["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        traffic[k] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "source": f"profiles/{tag}_ncu_summary.txt"}
    for k in ["chain_fast_kernel", "bass_wave_kernel", "gran_wave_kernel"]:
        if os.path.exists(f"gpurun_out/hot_lines_{tag}_{k}.txt"):
            shutil.copy(f"gpurun_out/hot_lines_{tag}_{k}.txt", f"profiles/{tag}_hot_lines_{k}.txt")
    warm = f"gpurun_out/ncu_{tag}_warm_dram.csv"
    if os.path.exists(warm):
        shutil.copy(warm, f"profiles/{tag}_ncu_warm_cache_dram.csv")
    open(f"profiles/{tag}_ncu_summary.txt", "w").write(out.getvalue())
    open(f"profiles/{tag}_sass_mix.txt", "w").write(mix.getvalue())
    json.dump(traffic, open("profiles/ncu_dram_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
