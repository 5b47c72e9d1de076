"""Host logic of the many-chain driver on CPU: sharding, the reference's checkpoint layout, resume bookkeeping, and the
multi-rank path (world_size 2, gloo) including the ensemble-moments all-reduce.  The GPU batch runner is replaced by an
oracle-backed runner, so no CUDA is needed here."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import TRAJECTORY_CASES
from gpu_helpers import oracle_setup
from mcmc_gpu_b200 import drivers
from oracle import crf_oracle as O

CASE = dict(TRAJECTORY_CASES["ragged_rf"])
N_CHAINS, N_ITER = 5, 12


def _oracle_runner(cs_fp, rf, beds, keys, iter0s, n_iter):
    cs, fp = cs_fp
    out = []
    for bed, key, it0 in zip(beds, keys, iter0s):
        seed = (key ^ (it0 * 0x9E3779B97F4A7C15)) % (2 ** 63)          # counter-addressed like the device RNG
        r = O.run_chain(cs, fp, bed, n_iter, np.random.default_rng(seed), np.random.default_rng(seed + 1))
        out.append((r["bed"], r["loss_mc"], r["loss_data"], r["loss"], r["steps"], r["resampled_times"], r["blocks"]))
    return out


def _inputs():
    g, cs, fp = oracle_setup(CASE)
    beds = [g["bed0"] + 0.5 * c for c in range(N_CHAINS)]
    seeds = [424200 + c for c in range(N_CHAINS)]
    return g, (cs, fp), beds, seeds


def test_shard_chains_partitions_exactly():
    for n in (0, 1, 7, 256, 4096, 4099):
        for world in (1, 2, 3, 8):
            parts = [drivers.shard_chains(n, world, r) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        drivers.shard_chains(4, 2, 2)


def test_checkpoint_layout_and_resume(tmp_path):
    g, cs_fp, beds, seeds = _inputs()
    res = drivers.largeScaleChain_mp(N_CHAINS, 4, cs_fp, None, beds, seeds, [N_ITER] * N_CHAINS, str(tmp_path),
                                     runner=_oracle_runner, verbose=False)
    assert len(res) == N_CHAINS and len(res[0]) == 7
    folder = tmp_path / "LargeScaleChain" / "424200"
    assert sorted(p.name for p in folder.iterdir()) == ["RNGState_RandField.txt", "RNGState_chain.txt", "bed_0k.npy",
                                                        "current_iter.txt", "results_0k.npz"]
    assert int(np.loadtxt(folder / "current_iter.txt")) == N_ITER
    st = json.loads((folder / "RNGState_chain.txt").read_text())
    assert st["bit_generator"] == drivers.RNG_KIND and st["iteration"] == N_ITER
    with np.load(folder / "results_0k.npz") as r:
        assert set(r.files) == {"loss_mc", "loss_data", "loss", "steps", "resampled_times", "blocks_used"}
        assert r["loss"].shape == (N_ITER,) and r["blocks_used"].shape == (N_ITER, 4)
    assert np.array_equal(np.load(folder / "bed_0k.npy"), res[0][0])
    # resume: continues from the saved bed and RNG position, appends like lsc_run_wrapper (:222-229)
    res2 = drivers.largeScaleChain_mp(N_CHAINS, 4, cs_fp, None, beds, seeds, [N_ITER] * N_CHAINS, str(tmp_path),
                                      runner=_oracle_runner, verbose=False)
    assert int(np.loadtxt(folder / "current_iter.txt")) == 2 * N_ITER
    with np.load(folder / "results_0k.npz") as r:
        assert r["loss"].shape == (2 * N_ITER,)
        cs = cs_fp[0]       # index 0 of the resumed run = loss recomputed from the saved bed (stale-ring quirk: not the tracked one)
        full = O.masked_loss(O.mass_conservation_residual(res[0][0], cs.surf, cs.velx, cs.vely, cs.dhdt, cs.smb, cs.resolution),
                             cs.mc_region_mask, cs.sigma_mc)[0]
        assert r["loss"][N_ITER] == full
        assert np.array_equal(r["resampled_times"], res[0][5] + res2[0][5])
    assert json.loads((folder / "RNGState_chain.txt").read_text())["iteration"] == 2 * N_ITER - 1
    # a numpy-PCG64 checkpoint written by the reference cannot be continued silently
    (folder / "RNGState_chain.txt").write_text(json.dumps({"bit_generator": "PCG64", "state": {}}))
    with pytest.raises(ValueError, match="different generator"):
        drivers.largeScaleChain_mp(1, 1, cs_fp, None, beds, seeds, [N_ITER], str(tmp_path), runner=_oracle_runner, verbose=False)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, cs_fp, beds, seeds = _inputs()
        res = drivers.largeScaleChain_mp(N_CHAINS, 1, cs_fp, None, beds, seeds, [N_ITER] * N_CHAINS, outdir,
                                         runner=_oracle_runner, verbose=False)
        mine = drivers.shard_chains(N_CHAINS, world, rank)
        assert len(res) == len(mine)
        ref = torch.as_tensor(g["bed0"])
        local = torch.stack([torch.as_tensor(r[0]) for r in res]) - ref
        s1, s2, n = local.sum(0), (local * local).sum(0), torch.tensor([float(len(res))], dtype=torch.float64)
        drivers.allreduce_moments(s1, s2, n)
        mean, var = drivers.moments_to_mean_var(ref, s1, s2, n)
        np.savez(os.path.join(outdir, f"ens_rank{rank}.npz"), mean=mean.numpy(), var=var.numpy(), n=n.numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_process(tmp_path):
    """N-rank run == single-process run re-partitioned (chains are independent); only the all-reduce needs a tolerance."""
    g, cs_fp, beds, seeds = _inputs()
    single = drivers.largeScaleChain_mp(N_CHAINS, 1, cs_fp, None, beds, seeds, [N_ITER] * N_CHAINS, str(tmp_path / "single"),
                                        runner=_oracle_runner, verbose=False)
    outdir = str(tmp_path / "dist")
    os.makedirs(outdir)
    mp.spawn(_rank_main, args=(2, _free_port(), outdir), nprocs=2, join=True)
    for c, seed in enumerate(seeds):
        a = np.load(tmp_path / "single" / "LargeScaleChain" / str(seed)[:6] / "bed_0k.npy")
        b = np.load(tmp_path / "dist" / "LargeScaleChain" / str(seed)[:6] / "bed_0k.npy")
        assert np.array_equal(a, b) and np.array_equal(a, single[c][0])
    stack = np.stack([r[0] for r in single])
    e0, e1 = np.load(os.path.join(outdir, "ens_rank0.npz")), np.load(os.path.join(outdir, "ens_rank1.npz"))
    assert e0["n"][0] == N_CHAINS and np.array_equal(e0["mean"], e1["mean"]) and np.array_equal(e0["var"], e1["var"])
    assert np.allclose(e0["mean"], stack.mean(0), rtol=0, atol=1e-9)
    assert np.allclose(e0["var"], stack.var(0), rtol=1e-9, atol=1e-12)
