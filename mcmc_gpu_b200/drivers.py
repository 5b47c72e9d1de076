"""Many-chain driver: the GPU replacement of the reference's process-parallel `largeScaleChain_mp`
(largeScaleChain_multiprocessing.py:19-240).

The reference starts one OS process per chain (`mp.Pool.starmap(lsc_run_wrapper, ...)`, :78-79) and each worker writes a
per-seed checkpoint folder (:200-238).  Here the chains of a rank are stepped concurrently by ONE kernel launch on that
rank's GPU; with `torch.distributed` initialised (one process per GPU, `torchrun`) the chains are sharded across ranks
with no communication while stepping, and `ensemble_mean_var` performs the only collective (an all-reduce of the per-cell
first and second moments).  The on-disk layout is the reference's, so `visualization.ipynb`-style consumers keep working:

    <output_path>/LargeScaleChain/<str(seed)[:6]>/bed_{k}k.npy, results_{k}k.npz, current_iter.txt,
                                                  RNGState_chain.txt, RNGState_RandField.txt

The RNG-state files hold the Philox (key, next iteration) pair instead of a numpy PCG64 state.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

import numpy as np

from . import MCMC

RNG_KIND = "gmc-philox4x32-10"


# ---------------------------------------------------------------------------------------------------------------------
# sharding
# ---------------------------------------------------------------------------------------------------------------------
def dist_info():
    """(rank, world_size) of the default torch.distributed group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_chains(n_chains: int, world_size: int, rank: int) -> list:
    """Contiguous, balanced partition of chain indices (sizes differ by at most one).  Chains carry their own seed, so a
    chain's trajectory does not depend on the rank or GPU count that runs it."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_chains, world_size)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


# ---------------------------------------------------------------------------------------------------------------------
# checkpoint files (reference layout: largeScaleChain_multiprocessing.py:135-180, 200-238)
# ---------------------------------------------------------------------------------------------------------------------
def seed_folder(output_path, seed) -> Path:
    return Path(output_path) / "LargeScaleChain" / f"{str(seed)[:6]}"


def ssc_seed_folder(output_path, lsc_seed, ssc_seed) -> Path:
    """Folder of one small-scale chain below its parent large-scale chain (largeScaleChain_multiprocessing.py:289)."""
    return Path(output_path) / "LargeScaleChain" / f"{str(lsc_seed)[:6]}" / "SmallScaleChain" / f"{str(ssc_seed)[:6]}"


def load_checkpoint(folder: Path):
    """Returns None or dict(cumulative_iters, bed, previous results, rng state) like lsc_run_wrapper's resume branch."""
    marker = folder / "current_iter.txt"
    if not marker.exists():
        return None
    cumulative = int(np.loadtxt(marker))
    k = int(cumulative / 1000)
    with np.load(folder / f"results_{k}k.npz") as r:
        prev = {name: r[name] for name in ("loss_mc", "loss_data", "loss", "steps", "resampled_times", "blocks_used")}
    state = None
    f = folder / "RNGState_chain.txt"
    if f.exists():
        state = json.loads(f.read_text())
        if state.get("bit_generator") != RNG_KIND:
            raise ValueError(f"{f} was written by a different generator ({state.get('bit_generator')}); a numpy PCG64 stream "
                             "cannot be continued on the device — restart the chain or resume it with the reference")
    return dict(cumulative_iters=cumulative, k=k, bed=np.load(folder / f"bed_{k}k.npy"), previous=prev, rng=state)


def save_checkpoint(folder: Path, result, previous, cumulative_before: int, n_iter: int, key: int, next_iteration: int):
    """Append `result` (the reference's 7-tuple) to `previous` and write the reference's files."""
    folder.mkdir(parents=True, exist_ok=True)
    bed, loss_mc, loss_data, loss, steps, resampled, blocks = result
    stale = []
    if previous is not None:
        p = previous["previous"]
        loss_mc = np.concatenate([p["loss_mc"], loss_mc])
        loss_data = np.concatenate([p["loss_data"], loss_data])
        loss = np.concatenate([p["loss"], loss])
        steps = np.concatenate([p["steps"], steps])
        resampled = p["resampled_times"] + resampled
        blocks = np.vstack([p["blocks_used"], blocks])
        stale = [folder / f"results_{previous['k']}k.npz"]
    cumulative = cumulative_before + n_iter
    label = f"{cumulative // 1000}k"
    state = json.dumps({"bit_generator": RNG_KIND, "key": int(key), "iteration": int(next_iteration)})
    (folder / "RNGState_chain.txt").write_text(state)
    (folder / "RNGState_RandField.txt").write_text(state)
    np.save(folder / f"bed_{label}.npy", bed)
    np.savez_compressed(folder / f"results_{label}.npz", loss_mc=loss_mc, loss_data=loss_data, loss=loss, steps=steps,
                        resampled_times=resampled, blocks_used=blocks)
    for f in stale:
        if f.exists() and f.name != f"results_{label}.npz":
            f.unlink()
    np.savetxt(folder / "current_iter.txt", [cumulative], fmt="%d")


# ---------------------------------------------------------------------------------------------------------------------
# batch runners
# ---------------------------------------------------------------------------------------------------------------------
def gpu_runner(chain_obj, rf, beds, keys, iter0s, n_iter, device=None):
    """Run len(beds) chains for n_iter iterations (n_iter-1 proposals, index 0 = initial state) on this rank's GPU.
    Chains resumed at different Philox iterations are grouped by their starting iteration."""
    out = [None] * len(beds)
    for it0 in sorted(set(iter0s)):
        sel = [i for i, v in enumerate(iter0s) if v == it0]
        batch = MCMC.ChainBatch(chain_obj, rf, np.stack([beds[i] for i in sel]), [keys[i] for i in sel], iter0=it0,
                                device=device, track_resampled=True)
        res = batch.advance_into(n_iter - 1)
        resampled = batch.resampled_times()
        for j, i in enumerate(sel):
            loss = np.array(res["loss"][j], dtype=np.float64)
            out[i] = (np.array(res["bed"][j]), loss.copy(), np.zeros(n_iter), loss, np.array(res["steps"][j], dtype=np.float64),
                      resampled[j], np.array(res["blocks"][j], dtype=np.float64))
        batch.close()
    return out


def largeScaleChain_mp(n_chains, n_workers, largeScaleChain, rf, initial_beds, rng_seeds, n_iters, output_path="./Data/output",
                       *, runner=None, device=None, save=True, verbose=True):
    """Drop-in for the reference's largeScaleChain_mp (same positional signature and result tuples).

    n_workers is accepted for compatibility and ignored: parallelism comes from the GPU (all chains of a rank in one
    launch) and from torch.distributed ranks.  Returns the list of 7-tuples of THIS rank's chains, in chain order (with one
    rank that is every chain, as in the reference).  `runner(chain, rf, beds, keys, iter0s, n_iter)` can replace the GPU
    batch runner (used by the CPU tests).
    """
    rank, world = dist_info()
    mine = shard_chains(n_chains, world, rank)
    runner = runner or (lambda *a: gpu_runner(*a, device=device))
    tic = time.time()
    results = {}
    # chains with equal run length advance together
    for n_iter in sorted({int(n_iters[i]) for i in mine}):
        group = [i for i in mine if int(n_iters[i]) == n_iter]
        beds, keys, iter0s, prevs, cums = [], [], [], [], []
        for i in group:
            folder = seed_folder(output_path, rng_seeds[i])
            prev = load_checkpoint(folder) if save else None
            key = MCMC.philox_key(rng_seeds[i], rng_seeds[i])
            it0 = 1
            bed = np.asarray(initial_beds[i], dtype=np.float64)
            if prev is not None:
                bed = prev["bed"]
                if prev["rng"] is not None:
                    key, it0 = int(prev["rng"]["key"]), int(prev["rng"]["iteration"])
            beds.append(bed); keys.append(key); iter0s.append(it0); prevs.append(prev)
            cums.append(0 if prev is None else prev["cumulative_iters"])
        outs = runner(largeScaleChain, rf, beds, keys, iter0s, n_iter)
        for i, res, prev, cum, key, it0 in zip(group, outs, prevs, cums, keys, iter0s):
            results[i] = res
            if save:
                save_checkpoint(seed_folder(output_path, rng_seeds[i]), res, prev, cum, n_iter, key, it0 + n_iter - 1)
    if verbose and rank == 0:
        print(f"Completed in {time.time() - tic:.2f} seconds")
    return [results[i] for i in mine]


def sgs_gpu_runner(chain_obj, _unused, beds, keys, iter0s, n_iter, device=None):
    """Run len(beds) small-scale chains for n_iter block re-simulations on this rank's GPU (kernel K6)."""
    out = [None] * len(beds)
    for it0 in sorted(set(iter0s)):
        sel = [i for i, v in enumerate(iter0s) if v == it0]
        batch = MCMC.SgsBatch(chain_obj, np.stack([beds[i] for i in sel]), [keys[i] for i in sel], iter0=it0, device=device)
        lc, st, bl = batch.advance(n_iter)
        last = batch.beds(with_trend=True)
        resampled = batch.resampled_times()
        for j, i in enumerate(sel):
            loss = np.array(lc[j], dtype=np.float64)
            out[i] = (np.array(last[j]), loss.copy(), np.zeros(n_iter), loss, np.array(st[j], dtype=np.float64), resampled[j],
                      np.array(bl[j], dtype=np.float64))
        batch.close()
    return out


def smallScaleChain_mp(n_chains, n_workers, smallScaleChain, initial_beds, ssc_rng_seeds, lsc_rng_seed, n_iters,
                       output_path="./Data/output", *, runner=None, device=None, save=True, verbose=True):
    """Drop-in for the reference's smallScaleChain_mp (largeScaleChain_multiprocessing.py:243-318): same positional
    signature, same 7-tuples, same files under <output>/LargeScaleChain/<lsc seed>/SmallScaleChain/<ssc seed>/.
    All chains of a rank advance in one kernel launch; ranks (torch.distributed) take contiguous shards of the chains."""
    rank, world = dist_info()
    mine = shard_chains(n_chains, world, rank)
    runner = runner or (lambda *a: sgs_gpu_runner(*a, device=device))
    tic = time.time()
    results = {}
    for n_iter in sorted({int(n_iters[i]) for i in mine}):
        group = [i for i in mine if int(n_iters[i]) == n_iter]
        beds, keys, iter0s, prevs, cums = [], [], [], [], []
        for i in group:
            folder = ssc_seed_folder(output_path, lsc_rng_seed, ssc_rng_seeds[i])
            prev = load_checkpoint(folder) if save else None
            key, it0 = MCMC.philox_key(ssc_rng_seeds[i]), 0
            bed = np.asarray(initial_beds[i], dtype=np.float64)
            if prev is not None:
                bed = prev["bed"]
                if prev["rng"] is not None:
                    key, it0 = int(prev["rng"]["key"]), int(prev["rng"]["iteration"])
            beds.append(bed); keys.append(key); iter0s.append(it0); prevs.append(prev)
            cums.append(0 if prev is None else prev["cumulative_iters"])
        outs = runner(smallScaleChain, None, beds, keys, iter0s, n_iter)
        for i, res, prev, cum, key, it0 in zip(group, outs, prevs, cums, keys, iter0s):
            results[i] = res
            if save:
                save_checkpoint(ssc_seed_folder(output_path, lsc_rng_seed, ssc_rng_seeds[i]), res, prev, cum, n_iter, key,
                                it0 + n_iter)
    if verbose and rank == 0:
        print(f"Completed in {time.time() - tic} seconds")
    return [results[i] for i in mine]


# ---------------------------------------------------------------------------------------------------------------------
# ensemble statistics: the one collective on the path
# ---------------------------------------------------------------------------------------------------------------------
def allreduce_moments(sum_, sumsq, count, group=None):
    """In-place SUM all-reduce of (sum, sumsq, count) over the ranks as ONE packed buffer of 2*H*W+1 doubles (one collective
    launch instead of three); tensors may be CUDA (NCCL) or CPU (gloo)."""
    rank, world = dist_info()
    if world > 1:
        import torch
        import torch.distributed as dist
        n = sum_.numel()
        packed = torch.empty(2 * n + 1, dtype=sum_.dtype, device=sum_.device)
        packed[:n] = sum_.reshape(-1)
        packed[n:2 * n] = sumsq.reshape(-1)
        packed[2 * n] = count.reshape(-1)[0]
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sum_.copy_(packed[:n].reshape(sum_.shape))
        sumsq.copy_(packed[n:2 * n].reshape(sumsq.shape))
        count.copy_(packed[2 * n:2 * n + 1].reshape(count.shape))
    return sum_, sumsq, count


def moments_to_mean_var(ref_bed, sum_, sumsq, count, ddof=0):
    """mean = ref + S1/n;  var = (S2 - S1^2/n) / (n - ddof)  (moments are taken about ref_bed for conditioning)."""
    n = count if np.isscalar(count) else count.reshape(())
    mean = ref_bed + sum_ / n
    var = (sumsq - sum_ * sum_ / n) / (n - ddof)
    return mean, var


def ensemble_mean_var(batch, ref_bed=None, ddof=0, group=None):
    """Posterior mean/variance per cell over ALL chains of ALL ranks from device-resident beds (ChainBatch).

    Local part: kernel K5 (gmc_ensemble_moments) reads this rank's C*H*W beds once; global part: one all-reduce of
    2*H*W+1 doubles.  Returns CUDA tensors (mean[H,W], var[H,W]) identical on every rank.
    """
    import torch
    dev = batch.dev
    ref = batch.bed[0].clone() if ref_bed is None else torch.as_tensor(np.ascontiguousarray(ref_bed, dtype=np.float64)).to(dev)
    if ref_bed is None:
        _, world = dist_info()
        if world > 1:                                  # every rank must subtract the SAME reference
            import torch.distributed as dist
            dist.broadcast(ref, src=0, group=group)
    s1 = torch.empty((batch.H, batch.W), dtype=torch.float64, device=dev)
    s2 = torch.empty_like(s1)
    batch.ctx.ensemble_moments(batch.bed, ref, s1, s2)
    n = torch.tensor([float(batch.C)], dtype=torch.float64, device=dev)
    allreduce_moments(s1, s2, n, group)
    return moments_to_mean_var(ref, s1, s2, n, ddof)
