#!/usr/bin/env python
"""bench.py — chain-steps/s of the many-chain large-scale MCMC step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--chains-total T]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A bench "step" advances every chain by --iters Metropolis iterations (one fused kernel launch per GPU) with the FULL
output contract of the reference's chain_crf.run (final bed, loss / accept / block caches, resampled_times).  Workload at
N=1 is BASELINE.json configs[1]: 256 chains on the synthetic 500x500 grid (SURVEY.md 8d recipe); at N>1 every GPU holds
the same number of chains (weak scaling, no data-path collective; the ensemble-moments all-reduce runs once after the
timed region) unless --chains-total T shards T chains over the ranks (strong scaling, drivers.shard_chains).  The line's
`targets` object holds the north-star configurations measured in the same run: BASELINE.json configs[2] (4096 chains at
500x500 sharded over the N GPUs) and configs[4] (2000x2000, 1024 chains over 8 GPUs = 128 per GPU).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the end-to-end pipeline keeps (steps in flight) x (chain ranges) x 2 CUDA streams busy; with the default 8 hardware queues
# unrelated streams share a queue and serialise behind each other's 40 ms kernels
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

METRIC = "chain-steps/sec (chains x iters), large-scale chain"
UNIT = "chain-steps/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
REF_RUNNER = os.path.join(ROOT, "baseline", "run_reference.py")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=256, help="chains per GPU (weak scaling)")
    ap.add_argument("--chains-total", type=int, default=0, help="total chains, sharded over the ranks (strong scaling)")
    ap.add_argument("--grid", type=int, default=500)
    ap.add_argument("--iters", type=int, default=1000, help="Metropolis iterations per chain per bench step")
    ap.add_argument("--cpu-iters", type=int, default=0, help="iterations per chain of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the reference's own torch path on the GPU")
    ap.add_argument("--no-targets", action="store_true", help="skip the north-star target configurations")
    ap.add_argument("--target-iters", type=int, default=200)
    ap.add_argument("--no-sgs", action="store_true", help="skip the small-scale SGS chain sample")
    ap.add_argument("--sgs-chains", type=int, default=512)
    ap.add_argument("--sgs-iters", type=int, default=20)
    ap.add_argument("--port-baseline", action="store_true", help="CPU arm = the numpy oracle port even when baseline/_ref exists")
    return ap.parse_args()


def chains_per_rank(a, world, rank):
    if a.chains_total:
        from mcmc_gpu_b200.drivers import shard_chains
        ids = shard_chains(a.chains_total, world, rank)
        return len(ids), (ids[0] if ids else 0), a.chains_total
    return a.chains, rank * a.chains, world * a.chains


def base_config(a, world):
    """The configuration both arms report (identical keys and values for the same command line)."""
    from mcmc_gpu_b200 import synthetic as syn
    total = a.chains_total or world * a.chains
    per = f"{a.chains_total} chains sharded over {world} GPU(s)" if a.chains_total else f"{a.chains} chains/GPU"
    return {"workload": f"large-scale chain (random-field proposal + mass-conservation loss), {per}, synthetic {a.grid}x{a.grid} grid",
            "grid": [a.grid, a.grid], "chains_per_gpu": (a.chains_total + world - 1) // world if a.chains_total else a.chains,
            "chains_total": total, "iters_per_step": a.iters, "blocks": list(syn.BLOCKS), "field_model": "Matern nu=0.9 spectral",
            "outputs": "bed, loss/accept/block caches, resampled_times (full chain_crf.run contract)",
            "l2": "state (bed+residual) %.2f GB per GPU >> 126 MB L2: inputs larger than L2, no flush needed"
                  % (2 * ((a.chains_total + world - 1) // world if a.chains_total else a.chains) * a.grid * a.grid * 8 / 1e9)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm.  Preferred: the UNMODIFIED reference driver largeScaleChain_mp from baseline/_ref in its own process
# (baseline/run_reference.py, kind "reference").  Fallback when baseline/_ref is absent: the numpy oracle port of the
# reference's per-chain loop, one process per core (kind "port").
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


def port_arm(grid, n_iter, steps, warmup, cores=None):
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


def run_reference(mode, *args, timeout=900):
    """Run baseline/run_reference.py in its own process; returns its JSON dict or {"unavailable": why}."""
    if not os.path.isdir(os.path.join(REF_DIR, "gstatsMCMC")):
        return {"unavailable": "baseline/_ref absent"}
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, REF_RUNNER, mode, *[str(x) for x in args]], capture_output=True, text=True, timeout=timeout, env=env)
        lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": f"exit {r.returncode}: {r.stderr.strip()[-200:]}"}
        return json.loads(lines[-1])
    except Exception as e:                                       # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


def cpu_arm(a, n_iter, steps, warmup):
    """-> (value, cores, seconds per step, kind, sample description)."""
    cores = os.cpu_count() or 1
    if not a.port_baseline:
        r = run_reference("cpu_mp", "--grid", a.grid, "--iters", n_iter, "--steps", steps, "--warmup", warmup)
        if "value" in r:
            return (r["value"], r["cores"], r["seconds_per_step"], "reference",
                    f"{r['cores']} chains x {n_iter} iterations per step through the unmodified reference's largeScaleChain_mp "
                    f"(largeScaleChain_multiprocessing.py:19, one mp.Pool worker per chain, pool start-up and checkpoint I/O included), "
                    f"same {a.grid}x{a.grid} grid")
        sys.stderr.write(f"bench.py: reference driver unavailable ({r.get('unavailable')}); timing the oracle port instead\n")
        why = f"; the reference driver was unavailable here: {str(r.get('unavailable'))[-160:]}"
    else:
        why = ""
    v, c, sps = port_arm(a.grid, n_iter, steps, warmup, cores)
    return (v, c, sps, "port", f"{c} chains x {n_iter} iterations per step, numpy port of chain_crf.run (oracle/crf_oracle.py), "
                               f"one process per core, same {a.grid}x{a.grid} grid{why}")


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n_iter = a.cpu_iters or (1000 if a.grid <= 500 else 100)
    val, cores, sps, kind, sample = cpu_arm(a, n_iter, max(a.steps, 1), a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "strong" if a.chains_total else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": base_config(a, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_sidecar(name, info=None):
    """DRAM traffic and pipe figures of a kernel from the committed `ncu --set full` summary (profiles/r2/<name>.ncu.json,
    written by profiles/ncu_summary.py).  Returns None when the file is missing or was captured for a different build of the
    kernel (registers / block size / shared memory differ from what the library reports now)."""
    p = os.path.join(ROOT, "profiles", "r2", f"{name}.ncu.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
    except Exception:
        return None
    if info is not None:
        for k_side, k_info in (("block_size", "threads"), ("registers_per_thread", "registers")):
            if k_info in info and k_side in d and int(d[k_side]) != int(info[k_info]):
                return None
    d["source"] = os.path.relpath(p, ROOT)
    return d


def algorithmic_bytes(bl, st, H, W):
    """SURVEY 8d U3 bytes of the recorded proposals: halo-tile read + old-residual read + accepted write-back."""
    ix, iy, bh, bw = (bl[..., k].astype(np.int64) for k in range(4))
    ch_h = np.minimum(H, ix + bh // 2) - np.maximum(0, ix - bh // 2)
    ch_w = np.minimum(W, iy + bw // 2) - np.maximum(0, iy - bw // 2)
    return int((8 * ((ch_h + 2) * (ch_w + 2) + ch_h * ch_w) + st.astype(np.int64) * 16 * ch_h * ch_w).sum())


def build_chain(MCMC, syn, H, W, quiet):
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
    return g, ch, rf


def device_initial_beds(torch, bed0, chain0, C, dev, amplitude=5.0):
    """[C,H,W] initial beds on the device: bed0 plus a smooth bump whose phases depend on the GLOBAL chain id (the same
    recipe as synthetic.chain_initial_beds, evaluated by torch in chunks of 128 chains so 4096 x 500^2 beds need no 8 GB
    host array and only a handful of kernel launches)."""
    H, W = bed0.shape
    ii = torch.arange(H, dtype=torch.float64, device=dev)[None, :, None] / 37.0
    jj = torch.arange(W, dtype=torch.float64, device=dev)[None, None, :] / 41.0
    b0 = torch.as_tensor(bed0).to(dev)
    out = torch.empty((C, H, W), dtype=torch.float64, device=dev)
    ph = np.stack([np.random.default_rng(10_000 + chain0 + c).uniform(0.0, 2.0 * np.pi, size=2) for c in range(C)]) if C else np.zeros((0, 2))
    amp = np.where(np.arange(C) + chain0 == 0, 0.0, amplitude)
    for lo in range(0, C, 128):
        hi = min(C, lo + 128)
        p = torch.as_tensor(ph[lo:hi]).to(dev)
        a = torch.as_tensor(amp[lo:hi]).to(dev)[:, None, None]
        out[lo:hi] = b0[None] + a * torch.sin(ii + p[:, 0, None, None]) * torch.cos(jj + p[:, 1, None, None])
    return out


def timed_launches(torch, dist, world, local, batch, n_it, steps, warmup):
    """W untimed + K timed device-resident launches; returns (per-step ms list, total ms = max over ranks)."""
    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
    for _ in range(warmup):
        batch.advance(n_it, want_caches=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    ev[0].record()
    for k in range(steps):
        batch.ctx.run(batch.bed, batch.mcres, batch.ssq, batch.seeds, batch.iteration, n_it, *batch._device_caches(n_it), 0,
                      batch.resampled, 4096)
        batch.iteration += n_it
        ev[k + 1].record()
    barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=batch.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return step_ms, total_ms


def target_run(torch, dist, MCMC, syn, quiet, world, rank, local, dev, H, W, C, chain0, n_it, steps, warmup, peak, label):
    """One north-star configuration, device-resident: C chains of this rank at HxW, `steps` launches of n_it iterations."""
    g, ch, rf = build_chain(MCMC, syn, H, W, quiet)
    beds = device_initial_beds(torch, g["bed0"], chain0, C, dev)
    batch = MCMC.ChainBatch(ch, rf, beds, [MCMC.philox_key(1000 + chain0 + c, 1000 + chain0 + c) for c in range(C)], device=dev,
                            track_resampled=True)
    del beds
    step_ms, total_ms = timed_launches(torch, dist, world, local, batch, n_it, steps, warmup)
    lc, st, bl = (t.cpu().numpy() for t in batch._device_caches(n_it))
    total_c = C
    if world > 1:
        t = torch.tensor([float(C)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total_c = int(t.item())
    value = total_c * n_it * steps / (total_ms * 1e-3)
    by = algorithmic_bytes(bl, st, H, W)
    ach = by / (step_ms[-1] * 1e-3) / 1e9
    # stencil at this shape (U1: read bed + write residual; U2: fused residual + masked loss, no write-back)
    loss_d = torch.empty(C, dtype=torch.float64, device=dev)
    scratch = torch.empty_like(batch.bed)
    ms_u1 = time_call(torch, lambda: batch.ctx.residual(batch.bed, scratch))
    ms_u2 = time_call(torch, lambda: batch.ctx.residual_loss(batch.bed, None, loss_d, None))
    cells = C * H * W
    out = {"workload": label, "grid": [H, W], "chains_per_gpu": C, "chains_total": total_c, "n_gpus": world, "iters_per_step": n_it,
           "steps": steps, "warmup": warmup, "value": value, "unit": UNIT, "ms_per_step": total_ms / steps,
           "per_gpu_value": value / world, "acceptance_rate": float(st.mean()),
           "state_gb_per_gpu": 2 * C * H * W * 8 / 1e9, "step_kernel": batch.ctx.step_kernel_info(C),
           "roofline_U3": {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                           "algorithmic_bytes_per_chain_step": by / (C * n_it)},
           "stencil": {"U1_frac": cells * 16 / (ms_u1 * 1e-3) / 1e9 / peak, "U1_ms": ms_u1,
                       "U2_frac": cells * 8 / (ms_u2 * 1e-3) / 1e9 / peak, "U2_ms": ms_u2}}
    batch.close()
    del batch, scratch
    ch._ctx = None
    torch.cuda.empty_cache()
    return out


def time_call(torch, fn, reps=10):
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


def sgs_sample(a, dev, fp64_peak):
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
    # the same through the public objects with HOST arrays on both sides: C initial beds up (pinned), n_it iterations, the
    # final beds and the loss / accept / block caches down - set-up of the batch (normal scores, residual) inside the timed region
    beds_host = torch.empty((C, H, W), dtype=torch.float64).pin_memory()
    beds_host.copy_(torch.as_tensor(bed)[None].expand(C, H, W))
    keys = [MCMC.philox_key(s) for s in range(C)]
    out_host = torch.empty((C, H, W), dtype=torch.float64).pin_memory()

    def e2e_once():
        b = MCMC.SgsBatch(ch, beds_host.numpy(), keys, device=dev)
        caches = b.advance(n_it)
        final = b.beds(out=out_host)
        b.close()
        return final, caches
    e2e_once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    final, caches = e2e_once()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": C * n_it / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": int(beds_host.numel() * 8),
           "d2h_bytes_per_step": int(final.nbytes + sum(x.nbytes for x in caches)),
           "what": "MCMC.SgsBatch(pinned host beds) -> advance(iters) -> beds(out=pinned) on the host, wall clock, batch set-up included"}
    del final, caches, beds_host, out_host
    out = {"workload": f"small-scale SGS chain, {C} chains, 300x300, blocks 5-19, 48 neighbours, radius 30 km, Matern nu=1.2259",
           "chain_steps_per_s": C * n_it / (ms * 1e-3), "kriged_nodes_per_s": nodes / (ms * 1e-3), "iters": n_it, "ms": ms,
           "acceptance_rate": float(st.mean()), "e2e": e2e}
    # FP64 work of the kriging solves: Gauss-Jordan on the augmented 49 x 51 system = ~n^2 (n+2) FMA per node (n = 49)
    n = 49
    flops = 2.0 * n * n * (n + 2) * nodes
    out["kriging_fp64_tflops"] = flops / (ms * 1e-3) / 1e12
    if fp64_peak:
        out["fp64_peak_tflops"] = fp64_peak
        out["fp64_frac"] = out["kriging_fp64_tflops"] / fp64_peak
    side = ncu_sidecar("sgs_run_kernel")
    if side:
        out["ncu"] = side
    if not a.no_cpu_baseline:
        r = run_reference("sgs", "--grid", 300, "--iters", 6)
        if "it_per_s_one_chain" in r:
            out["cpu_baseline"] = {"value": r["it_per_s_one_chain"], "unit": UNIT, "cores": 1, "kind": "reference",
                                   "sample": "chain_sgs.run (MCMC.py:1599) of the unmodified reference, one process, 6 iterations, same 300x300 configuration"}
        else:
            out["cpu_baseline"] = {"unavailable": r.get("unavailable")}
    return out


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
    from mcmc_gpu_b200 import hostio
    affinity = hostio.pin_to_gpu_numa(local)                     # CPU affinity / first-touch locality of the pinned buffers
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
    C, chain0, total_chains = chains_per_rank(a, world, rank)
    g, ch, rf = build_chain(MCMC, syn, H, W, quiet)
    seeds = [1000 + chain0 + c for c in range(C)]                   # global chain id -> seed: invariant to GPU count
    n_it = a.iters
    NBUF = int(os.environ.get("GMC_E2E_STEPS_IN_FLIGHT", "2"))      # e2e: steps in flight (multi-buffered device state)
    E2E_RANGES = int(os.environ.get("GMC_E2E_RANGES", "16"))        # chain ranges per step (two streams each)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory()
    # Coverage counts travel as int32 (the device type) or int16 (narrowed on the device: -128 MB per step, +one kernel per
    # chain range that has to squeeze in between the step CTAs).  Measured: one GPU 38.5 ms per step with int32 vs 41.4 ms
    # with int16 (compute bound: the extra kernel costs more than the bytes save); eight GPUs share one host whose link
    # bounds the step (profiles/README.md), so there the narrower type is used.
    counts_bits = int(os.environ.get("GMC_E2E_COUNTS", "16" if world >= 4 else "32"))
    counts_dtype = torch.int16 if counts_bits == 16 else torch.int32
    host_beds = pinned((C, H, W), torch.float64)
    host_beds.copy_(device_initial_beds(torch, g["bed0"], chain0, C, dev).cpu())
    outs = [{"bed": pinned((C, H, W), torch.float64), "loss": pinned((C, n_it + 1), torch.float64),
             "steps": pinned((C, n_it + 1), torch.uint8), "blocks": pinned((C, n_it + 1, 4), torch.int32),
             "resampled": pinned((C, H, W), counts_dtype)} for _ in range(NBUF)]
    keys = [MCMC.philox_key(s, s) for s in seeds]
    batches = [MCMC.ChainBatch(ch, rf, host_beds, keys, device=dev, track_resampled=True) for _ in range(NBUF)]
    batch = batches[0]
    ctx = batch.ctx
    info = ctx.step_kernel_info(C)
    peak, peak_src = measured_peak()

    # ---- (1) device-resident throughput: inputs already in HBM ---------------------------------------------------
    # pre-warm: a freshly started B200 needs about a second under load before clocks / memory settle (first-run numbers
    # were 20-35 % low without it); reported separately, NOT counted as warm-up steps.  Then exactly --warmup untimed steps.
    sampler = ClockSampler(local)
    sampler.start()                                  # started before the warm-up: no idle gap in front of the timed region
    n_pre = 0
    t_w0 = time.perf_counter()
    while time.perf_counter() - t_w0 < float(os.environ.get("GMC_BENCH_PREWARM_S", "1.5")):     # 0 for ncu launch lists
        batch.advance(n_it, want_caches=False)
        torch.cuda.synchronize()
        n_pre += 1
    prewarm_s = time.perf_counter() - t_w0
    l0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    step_ms, total_ms = timed_launches(torch, dist, world, local, batch, n_it, a.steps, a.warmup)
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count() - l0 - a.warmup
    value = total_chains * n_it * a.steps / (total_ms * 1e-3)
    clocks = sampler.summary(t_wall0, t_wall1)

    # algorithmic HBM bytes of the last timed launch, from the proposals it actually drew
    lc, st, bl = (t.cpu().numpy() for t in batch._device_caches(n_it))
    bytes_last = algorithmic_bytes(bl, st, H, W)
    acc_rate = float(st.mean())
    launch_ms = float(np.mean(step_ms))
    achieved = bytes_last / (step_ms[-1] * 1e-3) / 1e9
    side = ncu_sidecar("run_kernel", info)
    traffic = None
    if side and side.get("chain_steps"):
        traffic = (side["dram_bytes_read"] + side["dram_bytes_write"]) / side["chain_steps"] * C * n_it
    fp64_peak = ctx.fp64_peak_tflops()
    roofline = {"kernel": "run_kernel (fused K1 field synthesis + K4 Metropolis step)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": (side["source"] + " (dram__bytes_read.sum + dram__bytes_write.sum per chain-step x chain-steps per launch)")
                if traffic else "no ncu capture of this build committed (profiles/r2/run_kernel.ncu.json absent or stale)",
                "algorithmic_bytes_per_launch": float(bytes_last), "peak_source": peak_src,
                "algorithmic_bytes_per_chain_step": float(bytes_last) / (C * n_it), "launch_ms": launch_ms,
                "note": "U3 block-local formulation (the accepted-block resampled_times update, 8 B/cell, is not counted); this kernel is "
                        "FP64 latency / issue bound, not HBM bound (DESIGN.md)",
                "fp64_peak_tflops_measured": fp64_peak, "ncu": side}
    if side and fp64_peak and side.get("fp64_flops_per_chain_step"):
        roofline["fp64_tflops"] = side["fp64_flops_per_chain_step"] * C * n_it / (launch_ms * 1e-3) / 1e12
        roofline["fp64_frac"] = roofline["fp64_tflops"] / fp64_peak

    # ---- (2) the stencil metric: fused full-grid residual + masked loss (U2) and residual write (U1) ----------------
    loss_d = torch.empty(C, dtype=torch.float64, device=dev)
    scratch_res = torch.empty_like(batch.bed)
    ms_u2 = time_call(torch, lambda: ctx.residual_loss(batch.bed, None, loss_d, None))
    ms_u1 = time_call(torch, lambda: ctx.residual(batch.bed, scratch_res))
    ms_cp = time_call(torch, lambda: scratch_res.copy_(batch.bed))
    cells = C * H * W
    stencil = {"U2_residual_loss": {"GBps": cells * 8 / (ms_u2 * 1e-3) / 1e9, "bytes_per_cell": 8, "ms": ms_u2},
               "U1_residual": {"GBps": cells * 16 / (ms_u1 * 1e-3) / 1e9, "bytes_per_cell": 16, "ms": ms_u1},
               "torch_copy_same_arrays": {"GBps": cells * 16 / (ms_cp * 1e-3) / 1e9, "ms": ms_cp}, "peak": peak,
               "kernel": ctx.stencil_kernel_name(), "ncu": ncu_sidecar("residual_tma_kernel"), "ncu_U2": ncu_sidecar("residual_tma_kernel_U2")}
    for k in ("U2_residual_loss", "U1_residual", "torch_copy_same_arrays"):
        stencil[k]["frac"] = stencil[k]["GBps"] / peak
    del scratch_res

    # ---- (3) end to end through the public API: pinned host beds in, host results out, every step --------------------
    # Every step uploads the C initial beds from pinned host memory and downloads beds, caches and resampled_times; two
    # steps are in flight (double-buffered device state), so the upload / download of one overlaps the compute of the other.
    def e2e_loop(n_steps, overlap):
        pend = [None] * NBUF
        for k in range(n_steps):
            b = k % NBUF if overlap else 0
            if pend[b] is not None:
                pend[b].wait()
            pend[b] = ch.run_many(n_it + 1, rf, host_beds, seeds, as_arrays=True, batch=batches[b], out=outs[b], wait=False,
                                  pipeline_groups=E2E_RANGES if overlap else 16)
            if not overlap:
                pend[b].wait()
                pend[b] = None
        for p in pend:
            if p is not None:
                p.wait()

    def e2e_time(overlap):
        e2e_loop(2, overlap)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(a.steps, overlap)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    e2e_serial_ms = e2e_time(False)
    e2e_ms = e2e_time(True)
    e2e_val = total_chains * n_it * a.steps / (e2e_ms * 1e-3)
    h2d = host_beds.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in outs[0].values())
    hostbw = hostio.probe_host_bandwidth(torch, dev, world, dist if world > 1 else None)
    sampler.stop()

    # ---- (4) the one collective: ensemble mean/variance over all chains of all GPUs (outside the timed region) --------
    from mcmc_gpu_b200 import drivers
    mean, var = drivers.ensemble_mean_var(batch, g["bed0"])          # K5 + one packed all-reduce(2*H*W+1 doubles)
    torch.cuda.synchronize()
    ref_bed = torch.as_tensor(g["bed0"]).to(dev)
    ens = {"chains": total_chains, "mean_abs_shift_m": float((mean - ref_bed).abs().mean().item()),
           "mean_std_m": float(var.clamp_min(0).sqrt().mean().item())}
    for b in batches:
        b.close()
    del batches, batch, outs, host_beds, mean, var
    ch._ctx = None
    torch.cuda.empty_cache()

    # ---- (5) north-star target configurations, device-resident, in the same run ---------------------------------------
    targets = None
    if not a.no_targets:
        targets = {}
        try:
            from mcmc_gpu_b200.drivers import shard_chains
            ids = shard_chains(4096, world, rank)
            targets["config3_4096x500"] = target_run(
                torch, dist, MCMC, syn, quiet, world, rank, local, dev, 500, 500, len(ids), ids[0], a.target_iters, 3, 3, peak,
                f"BASELINE.json configs[2]: 4096 chains at 500x500 sharded over {world} GPU(s)")
            targets["config5_2000grid"] = target_run(
                torch, dist, MCMC, syn, quiet, world, rank, local, dev, 2000, 2000, 128, rank * 128, a.target_iters, 3, 3, peak,
                "BASELINE.json configs[4] per-GPU share: 128 chains per GPU at 2000x2000 (1024 chains on 8 GPUs)")
        except Exception as e:                                   # noqa: BLE001  never lose the headline line over a target
            targets["error"] = repr(e)[:300]

    # ---- (6) secondary hot path: small-scale SGS chain (BASELINE.json config 4), a short device-resident sample -----------
    sgs = None
    if rank == 0 and not a.no_sgs:
        try:
            sgs = sgs_sample(a, dev, fp64_peak)
        except Exception as e:                                   # never lose the headline line over the secondary sample
            sgs = {"error": repr(e)[:200]}

    cpu = ref_gpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_iter_cpu = a.cpu_iters or (1000 if a.grid <= 500 else 100)
        v, cores, _, kind, sample = cpu_arm(a, n_iter_cpu, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if rank == 0 and world == 1 and not a.no_reference_gpu:
        # A8: the reference's own GPU implementation (torch float32, one chain per process) on this box
        r = run_reference("gpu", "--grid", a.grid, "--iters", 300, "--procs", min(os.cpu_count() or 1, 16), timeout=600)
        ref_gpu = r if "it_per_s_one_chain" in r else {"unavailable": r.get("unavailable")}

    if rank == 0:
        cfg = base_config(a, world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "prewarm": {"steps": n_pre, "seconds": prewarm_s, "why": "clock / memory settle time of a fresh box, outside --warmup"},
                "ms_per_step": total_ms / a.steps, "step_ms": [round(x, 3) for x in step_ms], "higher_is_better": True,
                "scaling": "strong" if a.chains_total else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "details": {"acceptance_rate": acc_rate, "step_kernel": info, "cpu_affinity": affinity},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / a.steps, "steps_in_flight": NBUF, "chain_ranges_per_step": E2E_RANGES,
                        "coverage_counts_dtype": "int%d" % counts_bits,
                        "serial": {"value": total_chains * n_it * a.steps / (e2e_serial_ms * 1e-3), "ms_per_step": e2e_serial_ms / a.steps,
                                   "what": "one step at a time (upload, compute and download of a step finish before the next starts)"},
                        "host_bandwidth": hostbw,
                        "api": "chain_crf.run_many(pinned host beds -> host beds, loss/step/block caches, resampled_times), wait=False handles"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stencil": stencil, "ensemble": ens,
                "targets": targets, "sgs": sgs, "cpu_baseline": cpu, "reference_gpu": ref_gpu}
        if hostbw and hostbw.get("gbps_both_dirs_all_ranks"):
            need = (h2d + d2h) * world / (e2e_ms / a.steps * 1e-3) / 1e9
            line["e2e"]["host_gbps_used"] = need
            line["e2e"]["frac_of_host_ceiling"] = need / hostbw["gbps_both_dirs_all_ranks"]
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
