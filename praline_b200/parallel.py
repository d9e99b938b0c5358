"""Multi-GPU plumbing of the all-vs-all path: one process per GPU, pairs sharded by equal DP
cell count (engine.allpairs_tiles), one NCCL all-gather of the per-shard score slices.

The reference's only parallel strategy farms pickled tasks to worker processes
(praline/core/manager.py:248-463); the pair list shards with no data-path exchange, so the
only collective is the assembly of the condensed score vector that GuideTreeBuilder turns into
its distance matrix (praline/component/tree.py:137-147).

Layout (`ShardedCondensed`): rank r owns the condensed slots [cuts[r], cuts[r+1]).  The device
buffer is `world` slices of one common width (the longest slice, rounded to 16 bytes); slot s of
rank r lives at buf[s + shift[r]], shift[r] = r * width - cuts[r].  The DP kernels write their
scores straight into the rank's slice (they get the base pointer buf + shift[rank]), the
all-gather runs IN PLACE on that buffer (input = the rank's own slice of the output), and the
distance-matrix kernel reads the sliced layout directly (pgpu_tree_distance with cuts/shift): no
staging copy on either side of the collective.
"""
import numpy as np
import torch
import torch.distributed as dist


class ShardedCondensed(object):
    """The condensed all-vs-all score vector in rank slices (see the module docstring)."""

    def __init__(self, slot_cuts, device, dtype=torch.float32):
        self.cuts = [int(c) for c in slot_cuts]
        self.world = len(self.cuts) - 1
        sizes = [self.cuts[r + 1] - self.cuts[r] for r in range(self.world)]
        self.sizes = sizes
        self.width = (max(max(sizes), 1) + 3) & ~3
        self.shift = [r * self.width - self.cuts[r] for r in range(self.world)]
        self.n_slots = self.cuts[-1]
        self.buf = torch.empty(self.world * self.width, dtype=dtype, device=device)
        self._layout_dev = None

    def slice_of(self, rank):
        """The slots of `rank` as a contiguous view (slot cuts[rank] first)."""
        return self.buf[rank * self.width:rank * self.width + self.sizes[rank]]

    def where(self, slots):
        """Buffer positions of condensed slots (numpy int64 array in, array out)."""
        slots = np.asarray(slots, np.int64)
        r = np.searchsorted(np.asarray(self.cuts[1:], np.int64), slots, side="right")
        return slots + np.asarray(self.shift, np.int64)[np.minimum(r, self.world - 1)]

    def allgather(self, rank, group=None):
        """One in-place all-gather: afterwards every rank holds every slice."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf, self.buf[rank * self.width:(rank + 1) * self.width], group=group)
        return self

    def condensed(self):
        """The plain np.triu_indices-ordered vector (a copy unless there is one slice)."""
        if self.world == 1:
            return self.buf[:self.n_slots]
        return torch.cat([self.slice_of(r) for r in range(self.world)])

    def layout_dev(self):
        """(cuts, shift) as device int64 tensors for pgpu_tree_distance."""
        if self._layout_dev is None:
            dev = self.buf.device
            self._layout_dev = (torch.tensor(self.cuts, dtype=torch.int64, device=dev),
                                torch.tensor(self.shift, dtype=torch.int64, device=dev))
        return self._layout_dev


def allgather_condensed(out, slot_cuts, group=None):
    """Plain-vector form: every rank has filled out[slot_cuts[r]:slot_cuts[r+1]] of a contiguous
    condensed vector; after the call every rank holds the complete vector.  One broadcast per
    slice, in place (slices differ in length, so this is not an equal-sized all-gather; use
    ShardedCondensed for the one-collective layout).  Works for NCCL and for gloo on CPU tensors."""
    world = dist.get_world_size(group)
    if world == 1:
        return out
    for r in range(world):
        if slot_cuts[r + 1] > slot_cuts[r]:
            dist.broadcast(out[slot_cuts[r]:slot_cuts[r + 1]], src=dist.get_global_rank(group, r) if group else r,
                           group=group)
    return out


def scores_to_distance(cond, n):
    """Condensed scores -> the distance matrix of GuideTreeBuilder (tree.py:92-147), torch form used
    by the CPU (gloo) test; the device path is Engine.tree_distance (pgpu_tree_distance):
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


def allreduce_counts(counts, group=None):
    """Symmetric preprofile stage (Engine.preprofile_stage with shard=(rank, world)): the UNORDERED pair list
    is cut by DP cells, a rank's walks add into full-size count tables of every master (both orientations of
    a pair land in different masters' tables), and one sum all-reduce -- exact, the tables are int32 --
    hands every rank the complete tables (430 MB at BASELINE config 3: milliseconds on NVLink).  In place."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class SharedHostVector(object):
    """One pinned host vector shared by the ranks of a node (POSIX shared memory registered with
    the CUDA driver in every process): each rank copies ITS slice of the condensed vector device ->
    host over its own PCIe link, and the assembled vector is readable by every process without one
    rank pulling all of it through a single link.  Rank 0 creates the segment, the others attach
    (the name travels by a broadcast of a Python object)."""

    def __init__(self, n_elems, rank, world, dtype=np.float32, group=None):
        from multiprocessing import shared_memory
        self.rank, self.world = rank, world
        nbytes = max(int(n_elems) * np.dtype(dtype).itemsize, 16)
        name = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
            if rank != 0:
                self.shm = shared_memory.SharedMemory(name=name[0])
                try:        # Python < 3.13 registers attached segments with the tracker too: only the creator unlinks
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
        self.array = np.ndarray((int(n_elems),), dtype=dtype, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            self._registered = (int(rc) == 0)

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = None
        self.array = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass
