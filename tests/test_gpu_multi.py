"""Multi-GPU parity on hardware (SURVEY 8e's own test definition): a 2-rank NCCL run of the many-chain driver equals the
1-rank run chain for chain, bit for bit, and the C-ABI collective gmc_allreduce_moments (raw ncclComm_t) equals the
torch.distributed all-reduce.  Skipped on boxes with fewer than two GPUs; the host-side logic of the same path is covered
on CPU with gloo by tests/test_driver_dist.py."""
import ctypes
import os
import socket

import numpy as np
import pytest

from cases import TRAJECTORY_CASES

pytestmark = pytest.mark.gpu
CASE = dict(TRAJECTORY_CASES["ragged_rf"])
N_CHAINS, N_ITER = 7, 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _chain_and_inputs():
    from gpu_helpers import product_chain
    ch, rf, g = product_chain(CASE)
    beds = [g["bed0"] + 0.5 * c for c in range(N_CHAINS)]
    seeds = [515100 + c for c in range(N_CHAINS)]
    return ch, rf, g, beds, seeds


class _NcclId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


def _raw_nccl_comm(rank, world, dev):
    """A raw ncclComm_t next to torch's process group: rank 0 draws the unique id, torch.distributed carries it."""
    import torch
    import torch.distributed as dist
    path = None
    with open("/proc/self/maps") as f:
        for ln in f:
            if "libnccl.so" in ln:
                path = ln.split()[-1]
                break
    lib = ctypes.CDLL(path or "libnccl.so.2", mode=ctypes.RTLD_GLOBAL)
    uid = _NcclId()
    if rank == 0:
        assert lib.ncclGetUniqueId(ctypes.byref(uid)) == 0
    t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    ctypes.memmove(ctypes.byref(uid), raw, 128)
    comm = ctypes.c_void_p()
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _NcclId, ctypes.c_int]
    assert lib.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
    return lib, comm


def _rank_main(rank, world, port, outdir):
    import contextlib
    import io
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from mcmc_gpu_b200 import MCMC, _lib, drivers
        with contextlib.redirect_stdout(io.StringIO()):
            ch, rf, g, beds, seeds = _chain_and_inputs()
            res = drivers.largeScaleChain_mp(N_CHAINS, 1, ch, rf, beds, seeds, [N_ITER] * N_CHAINS, outdir, device=dev, verbose=False)
        mine = drivers.shard_chains(N_CHAINS, world, rank)
        assert len(res) == len(mine)
        # ensemble moments of this rank's final beds: torch.distributed path vs the C-ABI collective on a raw ncclComm_t
        with contextlib.redirect_stdout(io.StringIO()):
            batch = MCMC.ChainBatch(ch, rf, np.stack([r[0] for r in res]), [MCMC.philox_key(seeds[i]) for i in mine], device=dev)
        mean, var = drivers.ensemble_mean_var(batch, g["bed0"])
        ref = torch.as_tensor(np.ascontiguousarray(g["bed0"], dtype=np.float64)).to(dev)
        s1 = torch.empty_like(ref)
        s2 = torch.empty_like(ref)
        n = torch.tensor([float(batch.C)], dtype=torch.float64, device=dev)
        batch.ctx.ensemble_moments(batch.bed, ref, s1, s2)
        lib, comm = _raw_nccl_comm(rank, world, dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().gmc_allreduce_moments(batch.ctx._h, comm, s1.data_ptr(), s2.data_ptr(), n.data_ptr(), st))
        torch.cuda.synchronize()
        mean2, var2 = drivers.moments_to_mean_var(ref, s1, s2, n)
        lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        lib.ncclCommDestroy(comm)
        np.savez(os.path.join(outdir, f"ens_rank{rank}.npz"), mean=mean.cpu().numpy(), var=var.cpu().numpy(),
                 mean2=mean2.cpu().numpy(), var2=var2.cpu().numpy(), n=n.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpus_nccl_equal_one_gpu_bit_for_bit(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    from gpu_helpers import quiet
    from mcmc_gpu_b200 import drivers
    ch, rf, g, beds, seeds = quiet(_chain_and_inputs)
    single = quiet(drivers.largeScaleChain_mp, N_CHAINS, 1, ch, rf, beds, seeds, [N_ITER] * N_CHAINS, str(tmp_path / "single"),
                   verbose=False)
    outdir = str(tmp_path / "dist")
    os.makedirs(outdir)
    mp.spawn(_rank_main, args=(2, _free_port(), outdir), nprocs=2, join=True)
    for c, seed in enumerate(seeds):
        a = np.load(tmp_path / "single" / "LargeScaleChain" / str(seed)[:6] / "bed_0k.npy")
        b = np.load(tmp_path / "dist" / "LargeScaleChain" / str(seed)[:6] / "bed_0k.npy")
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) and np.array_equal(a, single[c][0])
        with np.load(tmp_path / "single" / "LargeScaleChain" / str(seed)[:6] / "results_0k.npz") as ra, \
                np.load(tmp_path / "dist" / "LargeScaleChain" / str(seed)[:6] / "results_0k.npz") as rb:
            for k in ("loss", "steps", "resampled_times", "blocks_used"):
                assert np.array_equal(ra[k], rb[k], equal_nan=True), k
    stack = np.stack([r[0] for r in single])
    e0, e1 = np.load(os.path.join(outdir, "ens_rank0.npz")), np.load(os.path.join(outdir, "ens_rank1.npz"))
    assert e0["n"][0] == N_CHAINS
    for k in ("mean", "var", "mean2", "var2"):
        assert np.array_equal(e0[k], e1[k]), k                     # every rank holds the same result
    # the C-ABI collective on the raw communicator == torch.distributed's all-reduce
    assert np.array_equal(e0["mean"], e0["mean2"]) and np.array_equal(e0["var"], e0["var2"])
    assert np.allclose(e0["mean"], stack.mean(0), rtol=0, atol=1e-9)
    assert np.allclose(e0["var"], stack.var(0), rtol=1e-9, atol=1e-12)


def test_bounded_wait_timeout_is_reported_not_ignored(monkeypatch):
    """a (chunk, chain) item that gives up waiting for its predecessor must surface as GmcError
    (GMC_ECUDA) instead of running on with stale state.  GMC_DEBUG_SPIN_LIMIT=0 makes every such wait give up at once."""
    import torch
    from gpu_helpers import product_chain, quiet
    from mcmc_gpu_b200 import MCMC
    from mcmc_gpu_b200._lib import GmcError
    monkeypatch.setenv("GMC_DEBUG_SPIN_LIMIT", "0")                 # read by gmc_create
    case = dict(TRAJECTORY_CASES["tutorial200"])
    ch, rf, g = quiet(product_chain, case)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    C = 2 * sms + 8                                                 # just more chains than resident CTAs: scheduled launch
    batch = MCMC.ChainBatch(ch, rf, np.stack([g["bed0"]] * C), list(range(C)))
    with pytest.raises(GmcError, match="bounded in-kernel wait"):
        for _ in range(4):                                          # every chunk boundary is a chance to time out
            batch.advance(64)
    # the flag is cleared by the report: a fresh context with the default limit runs normally
    monkeypatch.delenv("GMC_DEBUG_SPIN_LIMIT")
    ch._ctx = None
    ok = MCMC.ChainBatch(ch, rf, np.stack([g["bed0"]] * C), list(range(C)))
    lc, st, _ = ok.advance(64)
    assert np.isfinite(lc).all() and 0 < st.mean() < 1
