"""The one collective of the multi-GPU path: the periodic online -> target parameter broadcast.

BASELINE.json configs[4]: N independent env/replay/learner shards, one per GPU, exchange nothing
except the target-network refresh.  `broadcast_target` implements the "shared target" reading
(DESIGN.md §6): at every target-update boundary rank `src`'s ONLINE blob becomes every rank's TARGET
blob (ncclBroadcast over NVLink/NVSwitch when the tensors are CUDA tensors and the backend is nccl;
the same code runs on CPU tensors with gloo for the host-logic tests).  With world_size == 1 it
degenerates to the reference's `target_params = online_params` (dqn/agent.py:155-156).
"""

import torch


def broadcast_target(online: torch.Tensor, target: torch.Tensor, dist=None, src: int = 0) -> None:
  """target <- online(src) on every rank.  `dist` is `torch.distributed` (or None for one process)."""
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    target.copy_(online)
    return
  if dist.get_rank() == src:
    target.copy_(online)
  dist.broadcast(target, src=src)


def shard_seed(base_seed: int, rank: int) -> int:
  """Per-rank seed of the replay RandomState / synthetic contents (seed + rank)."""
  return int(base_seed) + int(rank)


def aggregate_throughput(steps_per_rank: int, world_size: int, max_ms_over_ranks: float) -> float:
  """Whole-job grad-steps/s: all ranks' steps over the slowest rank's device time."""
  return world_size * steps_per_rank / (max_ms_over_ranks / 1e3)
