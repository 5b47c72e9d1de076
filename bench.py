#!/usr/bin/env python
"""bench.py — chain-steps/s of the many-chain large-scale MCMC step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A bench "step" advances every chain by --iters Metropolis iterations (one fused kernel launch per GPU).  Workload at
N=1 is BASELINE.json configs[1]: 256 chains on the synthetic 500x500 grid (SURVEY.md §8d recipe); at N>1 every GPU
holds the same number of chains (weak scaling, no data-path collective; the ensemble-moments all-reduce runs once after
the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

METRIC = "chain-steps/sec (chains x iters), large-scale chain"
UNIT = "chain-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=256, help="chains per GPU")
    ap.add_argument("--grid", type=int, default=500)
    ap.add_argument("--iters", type=int, default=1000, help="Metropolis iterations per chain per bench step")
    ap.add_argument("--cpu-iters", type=int, default=0, help="iterations per chain of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sgs", action="store_true", help="skip the small-scale SGS chain sample")
    ap.add_argument("--sgs-chains", type=int, default=512)
    ap.add_argument("--sgs-iters", type=int, default=20)
    return ap.parse_args()


def workload_name(a):
    return f"large-scale chain (random-field proposal + mass-conservation loss), {a.chains} chains/GPU, synthetic {a.grid}x{a.grid} grid"


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle port of the reference's per-chain loop, one process per core (the reference's
# largeScaleChain_mp runs one chain per mp.Pool worker, largeScaleChain_multiprocessing.py:78-79)
# ---------------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(grid):
    from mcmc_gpu_b200 import synthetic as syn
    from oracle import crf_oracle as O
    os.environ["OMP_NUM_THREADS"] = "1"
    g = syn.make_grids(grid, grid)
    cs, fp = O.setup_from_grids(g, sigma_mc=syn.SIGMA_MC, logistic=syn.LOGISTIC, max_dist=syn.MAX_DIST, blocks=syn.BLOCKS)
    _CPU.update(g=g, cs=cs, fp=fp, O=O)


def _cpu_chain(args):
    seed, n_iter = args
    O, g = _CPU["O"], _CPU["g"]
    out = O.run_chain(_CPU["cs"], _CPU["fp"], g["bed0"], n_iter, np.random.default_rng(seed), np.random.default_rng(seed))
    return float(out["steps"].mean())


def cpu_arm(grid, n_iter, steps, warmup, cores=None):
    """Returns (chain-steps/s, cores, seconds per step).  Each step = `cores` chains x n_iter iterations."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")       # the parent may hold a CUDA context: never fork it
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(grid,)) as pool:
        pool.map(_cpu_chain, [(1, 3)] * cores)               # import + setup outside the timed region
        for w in range(warmup):
            pool.map(_cpu_chain, [(100 + c, max(n_iter // 10, 3)) for c in range(cores)])
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(_cpu_chain, [(1000 * (s + 1) + c, n_iter) for c in range(cores)])
        dt = time.perf_counter() - t0
    return cores * (n_iter - 1) * steps / dt, cores, dt / steps


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_iter = a.cpu_iters or (400 if a.grid <= 500 else 60)
    val, cores, sps = cpu_arm(a.grid, n_iter, max(a.steps, 1), a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "grid": [a.grid, a.grid]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{cores} chains x {n_iter} iterations per step, numpy port of chain_crf.run (oracle/crf_oracle.py), one process per core"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        # NVML from a sampling thread (a query takes well under a millisecond, so a 40 ms step gets several samples);
        # the nvidia-smi loop (>= 100 ms per row) is the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle, self.stop_flag = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index), False
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [("hw_slowdown", getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            sm_max = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        except Exception:
            sm_max = 0
        while not self.stop_flag:
            try:
                mask = int(reasons_fn(self.handle))
                row = [str(self.index), str(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), str(sm_max),
                       str(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0), hex(mask)]
                row += ["Active" if mask & b else "Not Active" for _, b in bits]
                self.rows.append((time.perf_counter(), row))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in ln.split(",")]))

    def stop(self):
        if getattr(self, "nvml", None):
            self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


def sgs_sample(a, dev):
    """chain-steps/s of the small-scale SGS chain (config 4: 512 chains, 300x300, blocks 5-19, 48 neighbours, 30 km
    radius, Matern nu=1.2259, normal-score transform + trend) on this rank's GPU."""
    import contextlib
    import io
    import torch
    from scipy.ndimage import gaussian_filter
    from sklearn.preprocessing import QuantileTransformer
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    H = W = 300
    g = syn.make_grids(H, W)
    bed = g["bed0"] + gaussian_filter(np.random.default_rng(99).standard_normal((H, W)), 2.0) * 30.0
    trend = gaussian_filter(bed, 10.0)
    nst = QuantileTransformer(n_quantiles=1000, output_distribution="normal", subsample=None, random_state=0).fit((bed - trend).reshape(-1, 1))
    with contextlib.redirect_stdout(io.StringIO()):
        ch = MCMC.chain_sgs(g["xx"], g["yy"], bed, g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"],
                            np.where(g["data_mask"] == 1, bed, np.nan), g["data_mask"], g["grounded_ice_mask"], g["resolution"])
        ch.set_update_region(True, g["highvel_mask"])
        ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
        ch.set_block_sizes(5, 20, 5, 20)
        ch.set_normal_transformation(nst, do_transform=True)
        ch.set_trend(trend, detrend_map=True)
        ch.set_variogram("Matern", 9932.5, 1.02, 0, isotropic=True, vario_smoothness=1.2259)
        ch.set_sgs_param(48, 30e3)
    C, n_it = a.sgs_chains, a.sgs_iters
    batch = MCMC.SgsBatch(ch, np.stack([bed] * C), [MCMC.philox_key(s) for s in range(C)], device=dev)
    batch.advance(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    lc, st, bl = batch.advance(n_it)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nodes = float((bl[..., 2].astype(np.float64) * bl[..., 3]).sum())
    batch.close()
    return {"workload": f"small-scale SGS chain, {C} chains, 300x300, blocks 5-19, 48 neighbours, radius 30 km, Matern nu=1.2259",
            "chain_steps_per_s": C * n_it / (ms * 1e-3), "kriged_nodes_per_s": nodes / (ms * 1e-3), "iters": n_it, "ms": ms,
            "acceptance_rate": float(st.mean())}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def gpu_main(a):
    import torch
    import torch.distributed as dist
    import contextlib
    import io
    from mcmc_gpu_b200 import MCMC, synthetic as syn

    def quiet(fn, *args, **kw):
        with contextlib.redirect_stdout(io.StringIO()):
            return fn(*args, **kw)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner / debug lines to stdout; stdout carries ONE JSON line, so send them to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    # ---- workload: tutorial configuration on the synthetic grid, through the public API -------------------------
    H = W = a.grid
    C = a.chains
    g = syn.make_grids(H, W)
    kw = syn.RF_KW
    rf = quiet(MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
               kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
    rf.set_block_sizes(*syn.BLOCKS)
    rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, g["resolution"])
    rf.set_generation_method(True)
    ch = quiet(MCMC.chain_crf, g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"],
               g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    quiet(ch.set_update_region, True, g["highvel_mask"])
    ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
    quiet(ch.set_update_type, "CRF_weight")
    ch.set_crf_data_weight(rf)
    seeds = [1000 + rank * C + c for c in range(C)]                 # global chain id -> seed: invariant to GPU count
    host_beds = torch.empty((C, H, W), dtype=torch.float64).pin_memory()
    host_beds.copy_(torch.as_tensor(syn.chain_initial_beds(g["bed0"], C)))
    n_it = a.iters
    out = {"bed": torch.empty((C, H, W), dtype=torch.float64).pin_memory(),
           "loss": torch.empty((C, n_it + 1), dtype=torch.float64).pin_memory(),
           "steps": torch.empty((C, n_it + 1), dtype=torch.uint8).pin_memory(),
           "blocks": torch.empty((C, n_it + 1, 4), dtype=torch.int32).pin_memory()}
    batch = MCMC.ChainBatch(ch, rf, host_beds, [MCMC.philox_key(s, s) for s in seeds], device=dev)
    ctx = batch.ctx
    info = ctx.step_kernel_info()

    # ---- (1) device-resident throughput: inputs already in HBM ---------------------------------------------------
    # warm-up: at least W (>= 3) untimed steps AND at least ~1.5 s of GPU work — a freshly started B200 needs about a
    # second under load before clocks/memory settle (first-run numbers were 20-35 % low with 3 x 43 ms of warm-up)
    sampler = ClockSampler(local)
    sampler.start()                                  # started before the warm-up: no idle gap in front of the timed region
    n_warm = 0
    t_w0 = time.perf_counter()
    while n_warm < max(a.warmup, 3) or time.perf_counter() - t_w0 < 1.5:
        batch.advance(n_it, want_caches=False)
        torch.cuda.synchronize()
        n_warm += 1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    barrier()
    batch.advance(n_it, want_caches=False)           # one more untimed step queued right behind the barrier
    l0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    ev[0].record()
    for k in range(a.steps):
        batch.ctx.run(batch.bed, batch.mcres, batch.ssq, batch.seeds, batch.iteration, n_it, *batch._device_caches(n_it), 0,
                      None, 4096)
        batch.iteration += n_it
        ev[k + 1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count() - l0
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * C * n_it * a.steps / (total_ms * 1e-3)
    clocks = sampler.summary(t_wall0, t_wall1)

    # algorithmic HBM bytes of the timed launches: halo-tile read + old-residual read + accepted write-back
    lc, st, bl = (t.cpu().numpy() for t in batch._device_caches(n_it))
    ix, iy, bh, bw = (bl[..., k].astype(np.int64) for k in range(4))
    ch_h = np.minimum(H, ix + bh // 2) - np.maximum(0, ix - bh // 2)
    ch_w = np.minimum(W, iy + bw // 2) - np.maximum(0, iy - bw // 2)
    bytes_last = (8 * ((ch_h + 2) * (ch_w + 2) + ch_h * ch_w) + st.astype(np.int64) * 16 * ch_h * ch_w).sum()
    acc_rate = float(st.mean())
    peak, peak_src = measured_peak()
    launch_ms = float(np.mean(step_ms))
    achieved = bytes_last / (step_ms[-1] * 1e-3) / 1e9
    # DRAM bytes per chain-step of this kernel from the committed ncu --set full capture (profiles/r1/run_kernel.ncu.txt:
    # 766.6 MB read + 250.3 MB written over 256 chains x 40 iterations), scaled to this launch's chain-steps
    ncu_traffic_per_chain_step = (766.582528e6 + 250.255872e6) / (256 * 40)
    roofline = {"kernel": "run_kernel (fused K1 field synthesis + K4 Metropolis step)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_per_chain_step * C * n_it,
                "traffic_source": "profiles/r1/run_kernel.ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum per chain-step x chain-steps per launch)",
                "algorithmic_bytes_per_launch": float(bytes_last), "peak_source": peak_src,
                "algorithmic_bytes_per_chain_step": float(bytes_last) / (C * n_it), "launch_ms": launch_ms,
                "note": "U3 block-local formulation; this kernel is FP64 latency / issue bound, not HBM bound (DESIGN.md)",
                "ncu": {"ipc_per_sm": 1.71, "issue_slots_pct": 47.3, "fp64_pipe_pct": 18.1, "dram_throughput_pct": 7.6,
                        "warps_per_sm": 16, "source": "profiles/r1/run_kernel.ncu.txt (ncu --set full, same kernel, 256 chains x 40 iterations)"}}

    # ---- (2) the stencil metric: fused full-grid residual + masked loss (U2) and residual write (U1) ----------------
    loss_d = torch.empty(C, dtype=torch.float64, device=dev)
    res_d = batch.mcres

    def time_call(fn, reps=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    scratch_res = torch.empty_like(batch.bed)
    ms_u2 = time_call(lambda: ctx.residual_loss(batch.bed, None, loss_d, None))
    ms_u1 = time_call(lambda: ctx.residual(batch.bed, scratch_res))
    cells = C * H * W
    stencil = {"U2_residual_loss": {"GBps": cells * 8 / (ms_u2 * 1e-3) / 1e9, "bytes_per_cell": 8, "ms": ms_u2},
               "U1_residual": {"GBps": cells * 16 / (ms_u1 * 1e-3) / 1e9, "bytes_per_cell": 16, "ms": ms_u1}, "peak": peak}
    for k in ("U2_residual_loss", "U1_residual"):
        stencil[k]["frac"] = stencil[k]["GBps"] / peak
    del scratch_res

    # ---- (3) end to end through the public API: pinned host beds in, host results out, every step --------------------
    for _ in range(2):
        ch.run_many(n_it + 1, rf, host_beds, seeds, as_arrays=True, batch=batch, out=out, track_resampled=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ch.run_many(n_it + 1, rf, host_beds, seeds, as_arrays=True, batch=batch, out=out, track_resampled=False)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_val = world * C * n_it * a.steps / (e2e_ms * 1e-3)
    h2d = host_beds.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in out.values())
    sampler.stop()

    # ---- (4) the one collective: ensemble mean/variance over all chains of all GPUs (outside the timed region) --------
    from mcmc_gpu_b200 import drivers
    mean, var = drivers.ensemble_mean_var(batch, g["bed0"])          # K5 + all-reduce(2*H*W+1 doubles) over the ranks
    torch.cuda.synchronize()
    ref_bed = torch.as_tensor(g["bed0"]).to(dev)
    ens = {"chains": world * C, "mean_abs_shift_m": float((mean - ref_bed).abs().mean().item()),
           "mean_std_m": float(var.clamp_min(0).sqrt().mean().item())}

    # ---- (5) secondary hot path: small-scale SGS chain (BASELINE.json config 4), a short device-resident sample -----------
    sgs = None
    if rank == 0 and not a.no_sgs:
        try:
            sgs = sgs_sample(a, dev)
        except Exception as e:                                   # never lose the headline line over the secondary sample
            sgs = {"error": repr(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_iter_cpu = a.cpu_iters or (1500 if a.grid <= 500 else 100)
        v, cores, _ = cpu_arm(a.grid, n_iter_cpu, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} chains x {n_iter_cpu} iterations, numpy port of chain_crf.run (oracle/crf_oracle.py), one process per core, same {a.grid}x{a.grid} grid"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": n_warm + 1,
                "ms_per_step": total_ms / a.steps, "step_ms": [round(x, 3) for x in step_ms], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(a), "grid": [H, W], "chains_per_gpu": C, "chains_total": world * C,
                           "iters_per_step": n_it, "blocks": list(syn.BLOCKS), "field_model": "Matern nu=0.9 spectral",
                           "l2": "state (bed+residual) %.2f GB per GPU >> 126 MB L2" % (2 * C * H * W * 8 / 1e9),
                           "acceptance_rate": acc_rate, "step_kernel": info},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / a.steps, "api": "chain_crf.run_many(pinned host beds -> host beds, loss/step/block caches)"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stencil": stencil, "ensemble": ens,
                "sgs": sgs, "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict):
    """The one JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    a = parse()
    # stdout carries exactly ONE JSON line: keep a private handle to it and point file descriptor 1 at stderr, so that
    # anything a library prints (NCCL's version banner is written to fd 1 from C) cannot interleave with it
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        reference_main(a)
    else:
        gpu_main(a)


if __name__ == "__main__":
    main()
