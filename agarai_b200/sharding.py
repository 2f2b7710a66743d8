"""Arena sharding over the GPUs of one box.

Arenas never interact, so the env path needs no collective: rank r of W owns the
contiguous block of global arena ids [r*n, (r+1)*n) and its engine handle is created with
`arena_id_base = r*n`.  The RNG is keyed by global arena id, so a rollout does not depend on W.
The only cross-rank operations are the timing reduction of bench.py (max over ranks) and, in
training, the gradient all-reduce of the policy network (agarai_b200/ppo.py).
"""
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    arenas_per_rank: int

    @property
    def arena_id_base(self) -> int:
        return self.rank * self.arenas_per_rank

    @property
    def total_arenas(self) -> int:
        return self.world * self.arenas_per_rank

    def global_ids(self):
        return range(self.arena_id_base, self.arena_id_base + self.arenas_per_rank)


def shard_from_env(arenas_per_rank: int) -> Shard:
    """Reads RANK / WORLD_SIZE as set by torchrun (1 process per GPU)."""
    return Shard(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(arenas_per_rank))


def split_total(total_arenas: int, world: int, rank: int) -> Shard:
    """Strong-scaling split of a fixed number of arenas (must divide evenly)."""
    if total_arenas % world:
        raise ValueError(f"{total_arenas} arenas do not divide over {world} ranks")
    return Shard(rank, world, total_arenas // world)


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's time: the number every multi-GPU throughput is computed from."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, seconds_local: float, device=None) -> float:
    """Whole-job units/s = all ranks' units / max-over-ranks time."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return world * units_per_rank / max_over_ranks(seconds_local, device)
