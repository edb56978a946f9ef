"""Multi-GPU host logic: shard images over ranks and gather per-rank statistics.

The path shards by image — the reference round-robins images over its workers
(``Benchmarker::get_next_worker_id`` benchmark-jpegxl/src/benchmark.rs:206-213, ``:449``) — so image ``i``
goes to rank ``i mod world`` and no pixel or coefficient ever crosses GPUs.  The only communication is one
reduction of a small statistics record per rank (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pins the calling process to the CPU cores of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated (first touch then places the buffers on that node): with one process per GPU and all ranks' input
    buffers on one socket, the end-to-end path of an 8-GPU box is limited by the inter-socket link instead of each
    GPU's own host link.  Best effort: returns what it did ({"numa_node": n, "cpus": k} or {"numa_node": None, ...})."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        pci = torch.cuda.get_device_properties(device_index)
        bus_id = f"{pci.pci_domain_id:04x}:{pci.pci_bus_id:02x}:{pci.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus_id}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:   # no sysfs / no permission: keep the default placement
        info["error"] = repr(e)[:80]
    return info


def shard_indices(num_images: int, rank: int, world: int) -> list[int]:
    """Indices of the images rank ``rank`` encodes: i mod world == rank (round robin, like the reference)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, num_images, world))


def distance_for_image(index: int, distances=(0.5, 1.0, 1.5, 2.0, 2.5, 3.0)) -> float:
    """BASELINE config 5: the distance sweep is assigned round-robin by GLOBAL image index."""
    return float(distances[index % len(distances)])


@dataclass
class ShardStats:
    images: int = 0
    pixels: int = 0
    codestream_bytes: int = 0
    device_ms: float = 0.0     # this rank's timed region
    kernel_launches: int = 0


@dataclass
class JobStats:
    images: int
    pixels: int
    codestream_bytes: int
    max_ms: float              # slowest rank: the job's time
    kernel_launches: int

    @property
    def mp_per_s(self) -> float:
        return self.pixels / 1e6 / (self.max_ms / 1e3) if self.max_ms > 0 else 0.0

    @property
    def bpp(self) -> float:
        return 8.0 * self.codestream_bytes / self.pixels if self.pixels else 0.0


def gather_stats(local: ShardStats, device: str = "cpu") -> JobStats:
    """SUM of the counters and MAX of the time over all ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return JobStats(local.images, local.pixels, local.codestream_bytes, local.device_ms, local.kernel_launches)
    sums = torch.tensor([local.images, local.pixels, local.codestream_bytes, local.kernel_launches], dtype=torch.int64,
                        device=device)
    mx = torch.tensor([local.device_ms], dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return JobStats(int(sums[0]), int(sums[1]), int(sums[2]), float(mx[0]), int(sums[3]))
