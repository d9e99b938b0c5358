"""Multi-GPU plumbing of the all-vs-all path: one process per GPU, pairs sharded by equal DP
cell count (engine.allpairs_tiles), one NCCL all-gather of the per-shard score slices.

The reference's only parallel strategy farms pickled tasks to worker processes
(praline/core/manager.py:248-463); the pair list shards with no data-path exchange, so the
only collective is the assembly of the condensed score vector that GuideTreeBuilder turns into
its distance matrix (praline/component/tree.py:137-147).
"""
import torch
import torch.distributed as dist


def allgather_condensed(out, slot_cuts, group=None):
    """Every rank has filled out[slot_cuts[r]:slot_cuts[r+1]]; after the call every rank holds
    the complete vector.  Slices are padded to the longest one so that one equal-sized
    all-gather does the exchange (works for NCCL on GPU and gloo on CPU tensors)."""
    world = dist.get_world_size(group)
    if world == 1:
        return out
    rank = dist.get_rank(group)
    sizes = [slot_cuts[r + 1] - slot_cuts[r] for r in range(world)]
    width = max(max(sizes), 1)
    send = torch.zeros(width, dtype=out.dtype, device=out.device)
    send[:sizes[rank]] = out[slot_cuts[rank]:slot_cuts[rank + 1]]
    recv = torch.empty(world * width, dtype=out.dtype, device=out.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    for r in range(world):
        if r != rank and sizes[r]:
            out[slot_cuts[r]:slot_cuts[r + 1]] = recv[r * width:r * width + sizes[r]]
    return out


def scores_to_distance(cond, n):
    """Condensed scores -> the distance matrix of GuideTreeBuilder (tree.py:92-147):
    d[i][j] = d[j][i] = score, d[i][i] = 0.0 (tree.py:132-133), dist = (-d) + d.max() in f32."""
    d = torch.zeros((n, n), dtype=torch.float32, device=cond.device)
    iu = torch.triu_indices(n, n, offset=1, device=cond.device)
    d[iu[0], iu[1]] = cond
    d[iu[1], iu[0]] = cond
    return (-d) + d.max()


def shard_masters(masters, lens, rank, world):
    """Contiguous shard of the (sorted, distinct) master ids with equal DP cells per rank.
    A master's work is len(master) * sum(len(slaves)), and every master sees (almost) the same
    slaves, so balancing on len(master) is enough."""
    import numpy as np
    masters = np.asarray(masters)
    w = np.cumsum(lens[masters].astype(np.float64))
    cuts = np.searchsorted(w, w[-1] * np.arange(world + 1) / world, side="left")
    cuts[0], cuts[-1] = 0, len(masters)
    return masters[cuts[rank]:cuts[rank + 1]], [int(c) for c in cuts]


def allgather_counts(local_counts, sizes, group=None):
    """Preprofile stage: every rank owns the whole [L x A] count tables of its masters
    (SURVEY.md 8e); one padded all-gather hands every rank all tables.  local_counts: 1-D tensor
    of this rank's tables; sizes[r]: number of elements rank r owns.  Returns the concatenation
    in rank order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_counts
    width = max(max(sizes), 1)
    send = torch.zeros(width, dtype=local_counts.dtype, device=local_counts.device)
    send[:local_counts.numel()] = local_counts
    recv = torch.empty(world * width, dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return torch.cat([recv[r * width:r * width + sizes[r]] for r in range(world)])
