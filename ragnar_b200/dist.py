"""One-process-per-GPU plumbing (torchrun): particle-range sharding and the
bootstrap of libragnar_cuda's own NCCL communicator.

The data path has exactly one exchange step: every rank reduces its shard to
`nbins` fp64 (+ u64 counts) and one all-reduce (sum) per result vector combines
them, issued by the library on its compute stream right behind the reduction
kernel (rgc_runtime.cu: allreduce_sum_* — a peer-store kernel over NVLink when the
ranks can map each other's exchange buffers, ncclAllReduce otherwise).
torch.distributed is used only to get the 128-byte NCCL unique id from rank 0 to
the other ranks.  The reference has no
multi-GPU path at all.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of [0, n_total): returns (offset, count) of `rank`.
    The first n_total % world ranks own one extra particle."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n_total), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def broadcast_bytes(dist, payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from `src` with torch.distributed
    (uint8 tensor on the GPU for the nccl backend, on the host for gloo)."""
    import torch

    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def install_communicator(cabi, dist) -> tuple[int, int]:
    """Create libragnar_cuda's NCCL communicator over the ranks of the initialised
    torch.distributed process group.  No-op (rank 0 of 1) when `dist` is None or
    the world has a single rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = cabi.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(dist, uid, cabi.COMM_ID_BYTES, src=0)
    cabi.comm_init(uid, rank, world)
    return rank, world


def allreduce_sum_host(dist, array: np.ndarray) -> np.ndarray:
    """Sum a small host array over ranks with torch.distributed (used by the CPU
    gloo tests of the sharding logic; the product path reduces on the device)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(array).copy())
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process (and the threads it creates afterwards: the library's staging pool)
    to the CPUs of the NUMA node the GPU hangs off, so pinned host buffers allocated from
    now on are node-local (first touch) and H2D copies do not cross the socket link.
    Returns what was done; a box without NUMA information is left alone."""
    import os

    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]  # 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except Exception as e:  # no NVML / sysfs: nothing to bind to
        info["error"] = f"{type(e).__name__}: {e}"
    return info
