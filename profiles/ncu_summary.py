#!/usr/bin/env python
"""Condense ncu reports into the text summaries kept under profiles/ (run in the build container, no GPU needed).
usage: ncu_summary.py <report.ncu-rep> [max kernels] [--json out.json --units N [--unit-name chain_steps] [--pick K]]

--json writes the machine-readable sidecar bench.py loads at run time (profiles/r2/<kernel>.ncu.json): DRAM bytes, pipe
figures and FP64 flops of kernel K of the report (default 0), with `units` = the work units that launch processed (e.g.
chains x iterations), so that per-unit traffic / flops can be scaled to any launch of the same build."""
import csv, io, json, subprocess, sys
argv = sys.argv[1:]
opt = {}
for flag in ("--json", "--units", "--unit-name", "--pick"):
    if flag in argv:
        i = argv.index(flag)
        opt[flag] = argv[i + 1]
        del argv[i:i + 2]
rep = argv[0]
kmax = int(argv[1]) if len(argv) > 1 else 3
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


def num(d, k, default=None):
    try:
        return float(d[k].replace(",", ""))
    except Exception:
        return default


if "--json" in opt:
    r = rows[2 + int(opt.get("--pick", 0))]
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    cyc = num(d, "sm__cycles_elapsed.max") or num(d, "smsp__cycles_elapsed.max") or 0.0
    flops = None
    keys = ["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
            "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"]
    if cyc and all(k in d for k in keys):
        flops = (num(d, keys[0]) + num(d, keys[1]) + 2.0 * num(d, keys[2])) * cyc
    n_units = float(opt.get("--units", 0)) or None
    name = opt.get("--unit-name", "chain_steps")
    side = {"kernel": d.get("Kernel Name", "?")[:120], "report": rep.split("/")[-1],
            "grid_size": int(num(d, "launch__grid_size", 0)), "block_size": int(num(d, "launch__block_size", 0)),
            "registers_per_thread": int(num(d, "launch__registers_per_thread", 0)),
            "duration_us": num(d, "gpu__time_duration.sum") * tscale.get(u.get("gpu__time_duration.sum", "us"), 1.0),
            "dram_bytes_read": num(d, "dram__bytes_read.sum") * scale.get(u.get("dram__bytes_read.sum", "byte"), 1.0),
            "dram_bytes_write": num(d, "dram__bytes_write.sum") * scale.get(u.get("dram__bytes_write.sum", "byte"), 1.0),
            "ipc_per_sm": num(d, "sm__inst_executed.avg.per_cycle_elapsed"),
            "issue_slots_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": num(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "icc_hit_rate_pct": num(d, "sm__icc_request_hit_rate.pct"), "fp64_flops": flops, name: n_units}
    if n_units:
        side["dram_bytes_per_" + name[:-1] if name.endswith("s") else "dram_bytes_per_" + name] = \
            (side["dram_bytes_read"] + side["dram_bytes_write"]) / n_units
        if flops:
            side["fp64_flops_per_" + (name[:-1] if name.endswith("s") else name)] = flops / n_units
    json.dump(side, open(opt["--json"], "w"), indent=1)
    print("wrote", opt["--json"])
