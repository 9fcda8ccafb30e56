"""Multi-GPU plumbing: one process per GPU, patches sharded in contiguous ranges, no data-path
collective; the only exchange is the (2C+1)-double sum of per-patch statistics.

Every rank draws ALL host indices from the single seeded stream and slices its range, so the
outputs are bit-identical for any world size.  Works over NCCL (device tensors) and gloo (CPU
tensors, used by the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import rng


@dataclass
class ShardPlan:
    n: int
    rank: int
    world: int

    @property
    def start(self) -> int:
        return rng.shard_range(self.n, self.rank, self.world)[0]

    @property
    def stop(self) -> int:
        return rng.shard_range(self.n, self.rank, self.world)[1]

    @property
    def count(self) -> int:
        return self.stop - self.start

    def take(self, arr):
        return arr[self.start:self.stop]


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch.distributed as dist
    rank, local, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


class numa_local:
    """Context manager: run the enclosed host code on the CPUs that are local to GPU `device` (sysfs
    `local_cpulist` of its PCI function), so that pinned staging buffers allocated inside land on the GPU's own
    NUMA node -- with 8 ranks pushing ~55 GB/s each, remote-socket staging memory halves the end-to-end rate.
    No-op when the topology cannot be read."""

    def __init__(self, device: int):
        self.device = device
        self.saved = None

    @staticmethod
    def _cpus(device: int):
        try:
            p = torch.cuda.get_device_properties(device)
            path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/local_cpulist"
            cpus = set()
            for part in open(path).read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            return cpus
        except Exception:
            return set()

    def __enter__(self):
        cpus = self._cpus(self.device)
        try:
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus and cpus != allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def local_stat_sums(mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """[N,C] per-patch means / stds of this rank -> f64 [2C+1] = (sum mean, sum std, N)."""
    c = mean.shape[1]
    out = torch.zeros(2 * c + 1, dtype=torch.float64, device=mean.device)
    out[:c] = mean.to(torch.float64).sum(dim=0)
    out[c:2 * c] = std.to(torch.float64).sum(dim=0)
    out[2 * c] = mean.shape[0]
    return out


def allreduce_stat_sums(sums: torch.Tensor) -> torch.Tensor:
    """SUM all-reduce of the 2C+1 doubles (88 bytes at C=5) when a process group exists."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def finish_stats(sums: torch.Tensor):
    """(avg_mean [C], avg_std [C], count): data_mean_std.py:45-46 over all ranks' patches."""
    s = sums.detach().cpu().numpy()
    c = (s.shape[0] - 1) // 2
    n = s[2 * c]
    return s[:c] / n, s[c:2 * c] / n, int(round(n))
