"""BASELINE config 3 at full size on this rank's GPUs: 10,000 synthetic 400-aa proteins.
(a) guide-tree stage on sequence tracks: all 49,995,000 unordered pairs, global, score per pair;
(b) preprofile stage: a sample of masters (40 per rank) against ALL other sequences, traced on the
    device into count tables, masters sharded by rank and the tables all-gathered (the full 10^8
    ordered pairs are (a)'s cells x2; reported extrapolated).
Prints JSON lines.  Single process per GPU (torchrun for N > 1 shards both stages)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from praline_b200 import get_engine, matrices, synth, parallel

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = get_engine(lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
S = matrices.blosum62()
t0 = time.perf_counter()
seqs = synth.family(3, n, 400)
batch = eng.batch(seqs)
S_dev = eng.dev(S)
t_gen = time.perf_counter() - t0
t0 = time.perf_counter()
plan = eng.allpairs_tiles(batch, (rank, world), paired=eng.wants_paired(S, -11.0, -1.0, 0, batch))
t_plan = time.perf_counter() - t0
out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=eng.device)
for rep in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.allpairs_scores(batch, S_dev, 27, [-11.0, -1.0], mode="global", shard=(rank, world), out=out, plan=plan, S_host=S)
    if world > 1: parallel.allgather_condensed(out, plan[3])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
t = torch.tensor([ms], device=eng.device)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
lens = batch.lens
cells = int((lens.sum() ** 2 - (lens ** 2).sum()) // 2)
if rank == 0:
    chk = float(out[:1000].sum().item())
    print(json.dumps({"stage": "C3 guide-tree all-vs-all (score only)", "n_seqs": n, "pairs": n * (n - 1) // 2, "cells": cells,
                      "n_gpus": world, "ms": float(t.item()), "gcups": cells / float(t.item()) / 1e6, "plan_s": t_plan, "gen_s": t_gen,
                      "checksum_first_1000": chk}))
# (b) preprofile stage on a sample of masters, sharded BY MASTER over the ranks: every rank
# traces its masters against all other sequences into its own count tables, one padded all-gather
# hands every rank all tables.
nm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
nm = max(nm, world) * (1 if world == 1 else 1)
all_masters = np.arange(nm * world if world > 1 else nm)
mine, cuts = parallel.shard_masters(all_masters, lens, rank, world)
masters = np.repeat(mine, n - 1)
slaves = np.concatenate([np.r_[0:i, i + 1:n] for i in mine]) if len(mine) else np.zeros(0, np.int64)
eng.preprofile_counts(batch, masters[:1000], slaves[:1000], S, [-11.0, -1.0])
for rep in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cnt, where, sc = eng.preprofile_counts(batch, masters, slaves, S, [-11.0, -1.0])
    if world > 1:
        local = torch.from_numpy(cnt).to(eng.device)
        sizes = [int(lens[all_masters[cuts[r]:cuts[r + 1]]].sum()) * 27 for r in range(world)]
        allc = parallel.allgather_counts(local, sizes)
        total_counts = int(allc.sum().item())
    else:
        total_counts = int(cnt.sum())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
td = torch.tensor([dt], dtype=torch.float64, device=eng.device)
pc = torch.tensor([float((lens[masters] * lens[slaves]).sum())], dtype=torch.float64, device=eng.device)
if world > 1:
    dist.all_reduce(td, op=dist.ReduceOp.MAX)
    dist.all_reduce(pc, op=dist.ReduceOp.SUM)
if rank == 0:
    dt, cells_pre = float(td.item()), float(pc.item())
    print(json.dumps({"stage": "C3 preprofile (traced -> count tables on device, masters sharded by rank, tables all-gathered)",
                      "n_gpus": world, "masters": int(len(all_masters)), "pairs": int(len(all_masters)) * (n - 1), "cells": cells_pre,
                      "wall_s": dt, "gcups": cells_pre / dt / 1e9, "full_stage_extrapolated_s": dt * n / len(all_masters),
                      "counts_sum": total_counts}))
# (c) the WHOLE preprofile stage, measured (argv[3] = "full" or "full-local"): all n masters sharded by rank,
# every master against all other sequences, chunked with nothing read back between chunks
# (Engine.preprofile_stage); count tables all-gathered at the end.
if len(sys.argv) > 3 and sys.argv[3].startswith("full"):
    pmode = "local" if sys.argv[3] == "full-local" else "global"
    nfull = int(sys.argv[4]) if len(sys.argv) > 4 else n          # masters taking part (default: all)
    allm = np.arange(nfull)
    mine, cuts = parallel.shard_masters(allm, lens, rank, world)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cnt_dev, where, cells_mine = eng.preprofile_stage(batch, S, [-11.0, -1.0], masters=mine, mode=pmode, iterations=2)
    t_host = time.perf_counter() - t0
    if world > 1:
        sizes = [int(lens[allm[cuts[r]:cuts[r + 1]]].sum()) * 27 for r in range(world)]
        allc = parallel.allgather_counts(cnt_dev, sizes)
    else:
        allc = cnt_dev
    total_counts = int(allc.sum(dtype=torch.int64).item())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    td = torch.tensor([dt], dtype=torch.float64, device=eng.device)
    pc = torch.tensor([float(cells_mine)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dist.all_reduce(pc, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"stage": "C3 preprofile stage, FULL (%s master-slave, count tables on device, all-gathered)" % pmode,
                          "n_gpus": world, "masters": int(nfull), "pairs": int(nfull) * (n - 1), "cells": float(pc.item()),
                          "wall_s": float(td.item()), "gcups": float(pc.item()) / float(td.item()) / 1e9,
                          "host_planning_s_rank0": t_host, "counts_sum": total_counts}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
