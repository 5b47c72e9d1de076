#!/usr/bin/env python
"""Condense ncu reports into the text summaries kept under profiles/ (run in the build container, no GPU needed).
usage: ncu_summary.py <report.ncu-rep> [max kernels]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:2 + kmax]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("=" * 110)
    print("kernel:", d.get("Kernel Name", "?")[:200])
    for k in KEYS:
        if k in d:
            print(f"  {k:88s} {d[k]:>16s} {u[k]}")
    st = sorted(((float(v.replace(",", "")), k[len(STALL):-len("_per_issue_active.ratio")]) for k, v in d.items()
                 if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
    print("  warp stall reasons (warps stalled per issue): " + ", ".join(f"{n} {x:.2f}" for x, n in st[:8]))
