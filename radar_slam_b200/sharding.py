"""Frame sharding across GPUs: frames are independent through the whole hot path, so a
sequence is split into contiguous blocks, one process per GPU, and the only collective is the
all-gather of the per-frame velocity rows [F, 8].  Works on any torch.distributed backend (NCCL on the
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def frame_block(rank: int, world: int, total_frames: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first total % world ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world) or total_frames < 0:
        raise ValueError("bad rank/world/total")
    base, rem = divmod(total_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_block(world: int, total_frames: int) -> int:
    return (total_frames + world - 1) // world


class VelocityGather:
    """Pre-allocated all-gather target.  `slot()` is this rank's [max_block, 8] slice: hand it to the velocity
    solve as its output so the collective needs no staging copy; `gather()` all-gathers in place and
    `assemble()` returns the [total_frames, 8] rows in frame order (ragged tails removed)."""

    def __init__(self, total_frames: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total = total_frames
        self.block = max_block(self.world, total_frames)
        self.buf = torch.zeros((self.world, self.block, 8), dtype=torch.float64, device=device)

    def slot(self) -> torch.Tensor:
        return self.buf[self.rank]

    def gather(self) -> torch.Tensor:
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf.view(-1), self.buf[self.rank].reshape(-1), group=self.group)
        return self.buf

    def assemble(self) -> torch.Tensor:
        parts = []
        for r in range(self.world):
            lo, hi = frame_block(r, self.world, self.total)
            parts.append(self.buf[r, : hi - lo])
        return torch.cat(parts, dim=0)


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (read from sysfs), so that the pinned
    host buffers it allocates afterwards are node-local and its H2D copies do not cross the socket interconnect.
    One process per GPU streams ~54 GB/s from host memory on the end-to-end path; with eight of them on a two-socket
    box the placement of those buffers decides whether the copies scale.  Best effort: returns what it did."""
    import os
    info = {"device": device_index, "numa_node": None, "cpus": None}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:           # no sysfs, no permission, odd topology: leave the affinity alone
        info["error"] = repr(e)
    return info
