"""Host-side plumbing of the end-to-end path: CPU/NUMA placement of a rank next to its GPU and a probe of what the host
can actually move over PCIe, so the end-to-end throughput can be stated as a fraction of a measured ceiling.

Nothing here computes on the hot path; it only places pinned buffers and measures copies (`cudaMemcpyAsync` through
torch's `copy_(non_blocking=True)`, one call per buffer)."""
from __future__ import annotations

import os
import time


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_info(local_index: int):
    """(pci bus id, numa node or None, cpus local to that node or empty set) of CUDA device `local_index`."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_index).pci_bus_id      # torch >= 2.3
        dom = getattr(torch.cuda.get_device_properties(local_index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(local_index).pci_device_id
        bdf = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
    except Exception:
        return None, None, set()
    base = f"/sys/bus/pci/devices/{bdf}"
    node = _read(base + "/numa_node")
    node = int(node) if node not in (None, "") else None
    cpus = _parse_cpulist(_read(base + "/local_cpulist"))
    if (not cpus) and node is not None and node >= 0:
        cpus = _parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist"))
    return bdf, node, cpus


def pin_to_gpu_numa(local_index: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off (when the container exposes that), BEFORE the
    pinned buffers are allocated and first touched, so they land in memory local to the GPU's PCIe root.  With several
    ranks on one node the allowed CPUs are additionally split evenly between the local ranks.  Returns what was done."""
    info = {"applied": False}
    try:
        bdf, node, cpus = gpu_numa_info(local_index)
        allowed = os.sched_getaffinity(0)
        info.update(pci=bdf, numa_node=node, allowed_cpus=len(allowed))
        want = (cpus & allowed) if cpus else set()
        if node is not None and node >= 0 and want and want != allowed:
            os.sched_setaffinity(0, want)
            info.update(applied=True, cpus=len(want))
        else:
            info["why_not"] = "no NUMA information for the GPU, or its local CPUs are all / none of the allowed set"
    except Exception as e:                                           # noqa: BLE001
        info["why_not"] = repr(e)[:120]
    return info


def probe_host_bandwidth(torch, dev, world=1, dist=None, mbytes=256, reps=6):
    """GB/s this host sustains for pinned H2D, D2H and both at once, measured on EVERY rank at the same time (barrier in,
    max time over ranks out) — the ceiling the end-to-end number of N ranks is compared with.  Plain per-buffer
    `cudaMemcpyAsync` copies on two streams."""
    n = mbytes * (1 << 20) // 8
    try:
        h_in = torch.empty(n, dtype=torch.float64).pin_memory()
        h_out = torch.empty(n, dtype=torch.float64).pin_memory()
        h_in.fill_(1.0)
        d_a = torch.empty(n, dtype=torch.float64, device=dev)
        d_b = torch.zeros(n, dtype=torch.float64, device=dev)
    except Exception as e:                                           # noqa: BLE001
        return {"error": repr(e)[:120]}
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(up, down):
        def once():
            if up:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        once()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return (int(up) + int(down)) * n * 8 * reps * world / dt / 1e9
    out = {"ranks": world, "mbytes_per_copy": mbytes,
           "gbps_h2d_all_ranks": run(True, False), "gbps_d2h_all_ranks": run(False, True),
           "gbps_both_dirs_all_ranks": run(True, True),
           "what": "pinned-memory copies issued by all ranks at once (one cudaMemcpyAsync per buffer, two streams), aggregate over ranks"}
    return out
