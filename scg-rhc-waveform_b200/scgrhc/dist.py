"""Multi-GPU plumbing of the hot path: one process per GPU under ``torchrun``, records sharded by
contiguous blocks (the per-record loop of recordutil.py:131-132 is embarrassingly parallel), NO data-path collective.

The only exchanges are
  * ``engine.allreduce_minmax``  one MIN all-reduce of {min, -max} (4 doubles) when ``use_global_min_max``
                                 (recordutil.py:152-169,185-189) — exact and order independent, so bit-identical to 1 GPU;
  * ``exchange_counts``          one all-gather of the per-rank kept-window counts: every rank learns its offset in the
                                 cohort-wide ordered list (records are sharded in order, so rank order == list order);
  * ``gather_rows``              rows of per-window tensors to rank 0 (valid / test windows for the pickles of
                                 recordutil.py:202-209; the train windows too when they fit one GPU).
torch.distributed supplies the transport (NCCL over NVLink between GPUs; gloo in the CPU tests, where collectives run
on host copies).
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


@dataclass
class Shard:
  rank: int = 0
  world: int = 1
  group: Optional[object] = None
  offset: int = 0           # position of this rank's first kept window in the cohort-wide ordered list
  total: int = 0            # kept windows over all ranks
  counts: tuple = ()        # kept windows per rank

  @property
  def active(self):
    return self.world > 1


def current(group=None):
  """(rank, world) of the running job; (0, 1) outside torch.distributed."""
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(group), dist.get_world_size(group)
  return 0, 1


def _host_collectives(group=None):
  """gloo moves host memory: collectives then run on CPU copies of the (small) tensors."""
  return dist.get_backend(group) != 'nccl'


def exchange_counts(n_local, device, group=None):
  """All-gather of one int64 per rank -> Shard with this rank's offset in the global ordered kept list."""
  rank, world = current(group)
  if world == 1:
    return Shard(0, 1, group, 0, int(n_local), (int(n_local),))
  dev = torch.device('cpu') if _host_collectives(group) else device
  mine = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
  every = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(world)]
  dist.all_gather(every, mine, group=group)
  counts = tuple(int(v.item()) for v in every)
  return Shard(rank, world, group, sum(counts[:rank]), sum(counts), counts)


def broadcast_index(idx, device, group=None, src=0):
  """Rank ``src``'s int64 index vector on every rank (the unseeded train/valid/test permutation of
  recordutil.py:191-192 must be ONE draw for the whole job)."""
  rank, world = current(group)
  if world == 1:
    return idx
  dev = torch.device('cpu') if _host_collectives(group) else device
  n = torch.tensor([idx.numel() if rank == src else 0], dtype=torch.int64, device=dev)
  dist.broadcast(n, src, group=group)
  buf = idx.to(dev, torch.int64).contiguous() if rank == src else torch.empty(int(n.item()), dtype=torch.int64, device=dev)
  dist.broadcast(buf, src, group=group)
  return buf.cpu()


def gather_rows(t, counts, group=None, dst=0):
  """Concatenation over ranks (rank order) of ``t`` (n_r, ...) on rank ``dst``; None elsewhere.  ``counts[r]`` = rows of
  rank r (known to every rank); shards travel padded to the longest one (device buffers with NCCL, host copies with
  gloo)."""
  rank, world = current(group)
  if world == 1:
    return t
  host = _host_collectives(group)
  src = t.cpu().contiguous() if host else t.contiguous()
  m = max(counts)
  tail = tuple(src.shape[1:])
  pad = torch.zeros((m,) + tail, dtype=src.dtype, device=src.device)
  pad[:src.shape[0]] = src
  out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
  dist.gather(pad, out, dst=dst, group=group)
  if rank != dst:
    return None
  return torch.cat([out[r][:counts[r]] for r in range(world)]).to(t.device)


def barrier(group=None):
  if current(group)[1] > 1:
    dist.barrier(group=group)
