"""world_size-2 gloo test of the shard + all-gather assembly (CPU tensors, no kernels)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from praline_b200 import engine as E, parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _B(object):
    pass


class _FakeEngine(E.Engine):
    def __init__(self):
        self.nw = 8
        self.k_set = [1, 2, 3, 4, 6, 8, 10, 12, 13, 14, 16, 20, 24, 32]
        self.pin = False


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _B()
    b.n = n
    b.lens = np.random.default_rng(1).integers(30, 300, n).astype(np.int64)
    eng = _FakeEngine()
    by_k, (lo, hi), cells, cuts, _, _ = eng.allpairs_tiles(b, (rank, world), tile=16)
    n_pairs = n * (n - 1) // 2
    out = torch.full((n_pairs,), float("nan"))
    # stand-in for the kernel: every slot of this shard gets a value derived from its pair
    pi, pj = synth.all_pairs(n)
    truth = torch.from_numpy((pi * 1000 + pj).astype(np.float32))
    for K, tiles in by_k.items():
        for t in tiles:
            a = int(t["out_base"])
            e = a + int(t["stream_end"] - t["stream_begin"])
            out[a:e] = truth[a:e]
    parallel.allgather_condensed(out, cuts)
    ok = bool(torch.equal(out, truth))
    # the sliced layout the GPU path uses: every rank fills its slice through the shift the kernels get
    # as a pointer offset, ONE in-place all-gather assembles it
    sc = parallel.ShardedCondensed(cuts, torch.device("cpu"))
    sc.buf.fill_(float("nan"))
    sc.buf[lo + sc.shift[rank]:hi + sc.shift[rank]] = truth[lo:hi]
    sc.allgather(rank)
    ok = ok and bool(torch.equal(sc.condensed(), truth))
    slots = np.arange(n_pairs)
    ok = ok and bool(np.array_equal(sc.buf.numpy()[sc.where(slots)], truth.numpy()))
    ok = ok and sc.width % 4 == 0 and sc.cuts == [int(c) for c in cuts]
    # the node-shared pinned host vector of the end-to-end path (each rank writes its own slice)
    hv = parallel.SharedHostVector(n_pairs, rank, world)
    hv.tensor[lo:hi] = truth[lo:hi]
    dist.barrier()
    ok = ok and bool(torch.equal(hv.tensor, truth))
    dist.barrier()
    hv.close()
    d = parallel.scores_to_distance(out, n)
    ok = ok and bool(d[3, 5] == d[5, 3]) and bool(d.min() == 0)
    # preprofile stage: masters sharded by rank, count tables all-gathered
    mine, cuts = parallel.shard_masters(np.arange(n), b.lens, rank, world)
    A = 5
    local = torch.cat([torch.full((int(b.lens[m]) * A,), float(m)) for m in mine]) if len(mine) else torch.zeros(0)
    sizes = [int(b.lens[cuts[r]:cuts[r + 1]].sum()) * A for r in range(world)]
    allc = parallel.allgather_counts(local, sizes)
    want = torch.cat([torch.full((int(b.lens[m]) * A,), float(m)) for m in range(n)])
    ok = ok and bool(torch.equal(allc, want)) and cuts[0] == 0 and cuts[-1] == n
    # symmetric preprofile stage: pairs sharded, every rank adds into full-size tables, one sum all-reduce
    part = torch.zeros(int(b.lens.sum()) * A, dtype=torch.int32)
    part[rank::world] = rank + 1
    total = parallel.allreduce_counts(part.clone())
    want_sum = torch.zeros_like(part)
    for r in range(world):
        want_sum[r::world] = r + 1
    ok = ok and bool(torch.equal(total, want_sum))
    q.put((rank, ok, cells))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_allgather_world2():
    world, n = 2, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    cells = sorted(c for _, _, c in res)
    assert cells[0] > 0.6 * cells[1]      # shards are balanced by DP cells
