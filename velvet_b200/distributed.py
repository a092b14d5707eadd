"""Multi-GPU plumbing for the batched-independent-cloths mode (north_star: "batched independent cloth instances shard
with no communication").  One process per GPU (torchrun); the data path has NO collective -- torch.distributed is used
only to agree on the timing (barrier + max over ranks).  Works with backend "nccl" on GPUs and "gloo" on CPU (tests).
"""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_instances(num_instances: int, world: int, rank: int) -> range:
    """Contiguous block of independent cloth instances owned by `rank` (instance k -> rank k // ceil(n / world));
    blocks differ in size by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(num_instances, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def instance_model_height(k: int) -> float:
    """instance k is placed at T(0, 1.5 + 0.01 * (k mod 32), 1) so instances differ."""
    return 1.5 + 0.01 * (k % 32)


class Group:
    """Thin wrapper: init / barrier / max-reduce of a python float.  No-op when WORLD_SIZE == 1."""

    def __init__(self, backend: str, device=None):
        self.rank, self.world, self.local_rank = env_rank_world()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend == "nccl":
                import torch
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, value: float) -> float:
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        if t.is_cuda:
            t = t.float()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value: float) -> float:
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        if t.is_cuda:
            t = t.float()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_all(self, obj) -> list:
        """Every rank's picklable `obj`, rank-ordered, on every rank (bookkeeping only: checksums, counts)."""
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, group: Group) -> float:
    """Whole-job throughput: units processed by all ranks / the slowest rank's time."""
    return group.sum(units_this_rank) / group.max(seconds_this_rank)
