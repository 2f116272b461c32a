"""Multi-GPU plumbing: worlds are independent (no cross-world term anywhere in
do_simulation, dart_env.py:158-175), so they shard as contiguous blocks, one process per GPU,
with ZERO communication during stepping.  The single optional collective is an all-gather of the
observation batch when the caller wants one flat tensor (replaces the mp.Array shared-memory
observation buffer of gym/vector/async_vector_env.py:90-94).

RNG streams are keyed by GLOBAL world id (world_offset + local index), so a world's trajectory
does not depend on how many GPUs the batch is sharded over.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_worlds(total_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of global world ids owned by `rank`: (offset, count).  The first
    total % world_size ranks get one extra world."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d not in [0, %d)" % (rank, world_size))
    base, rem = divmod(total_worlds, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_batch(local: torch.Tensor, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """All-gather equally sized per-rank batches [n_local, k] into [world_size * n_local, k], rows
    ordered by global world id.  NCCL over NVLink on GPUs (all_gather_into_tensor, one collective);
    falls back to list all_gather for backends without the tensor variant (gloo, CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    if out is None:
        out = torch.empty((ws * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        parts = list(out.chunk(ws, dim=0))
        dist.all_gather(parts, local.contiguous(), group=group)
    return out


class FusedObsGather:
    """The observation all-gather INSIDE the step kernel: every rank's kernel stores its observation rows straight into
    all ranks' gather buffers over NVLink (torch symmetric memory gives each rank the peers' buffer addresses), so the
    transfer overlaps the stepping tile by tile and what remains after the kernel is one device-side barrier.

        g = FusedObsGather(env)            # under torchrun, NCCL process group initialised
        obs_all = g.step(actions)          # [world_size * n_local, n_obs], rows ordered by global world id

    Raises RuntimeError where symmetric memory is unavailable (then use gather_batch: NCCL all_gather_into_tensor)."""

    def __init__(self, env, group=None):
        import torch.distributed._symmetric_memory as symm
        self.env, self.eng = env, env.engine
        group = group or dist.group.WORLD
        self.ws, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n, no = self.eng.n, self.eng.n_obs
        # two buffers, alternated per step: a rank may already be storing step k+1 into its peers while a peer's consumer of
        # step k is still reading (it cannot reach step k+2 before every rank has passed the barrier of step k+1)
        self.buf = symm.empty((2, self.ws * n, no), dtype=torch.float32, device=self.eng.device)
        self.hdl = symm.rendezvous(self.buf, group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.k = 0

    def step(self, actions: torch.Tensor):
        env = self.env
        n, no = self.eng.n, self.eng.n_obs
        half = self.k & 1
        self.eng.set_obs_peers(self.ptrs, half * self.ws * n * no + self.rank * n * no)
        self.eng.step(actions, env._obs, env._rew, env._done, env.auto_reset)
        self.hdl.barrier()                 # every rank's step kernel (and with it its peer stores) has completed
        self.k += 1
        return self.buf[half]

    def close(self):
        self.eng.set_obs_peers([], 0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing convention of bench.py: a multi-GPU step takes as long as its slowest rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def make_sharded(env_id: str, total_worlds: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                 local_rank: Optional[int] = None, **kw):
    """One DartEnv per rank over its shard of `total_worlds` (call under torchrun)."""
    import os

    from .envs import make
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size
    local_rank = int(os.environ.get("LOCAL_RANK", rank)) if local_rank is None else local_rank
    off, cnt = shard_worlds(total_worlds, rank, world_size)
    kw.setdefault("batched", True)
    return make(env_id, num_envs=cnt, device=local_rank, world_offset=off, **kw)
