"""Symmetric preprofile stage (Engine.preprofile_stage -> allpairs_dual: one traced fill per unordered pair,
two walks) on n synthetic 400-aa proteins of the config-3 family; prints wall/device time, GCUPS in
reference cells (ordered pairs) and in filled cells, and checks a sample of masters against the per-master path.

    python tools/run_dual.py [n_seqs=3000] [length=400] [check=1]
Env: PGPU_TB_GIB (traceback words per stream buffer), PGPU_DUAL_STREAMS, PGPU_TILE_DUAL."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
check = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = get_engine(0)
S = matrices.blosum62()
seqs = synth.family(3, n, L)
batch = eng.batch(seqs)
gaps = [-11.0, -1.0]
eng.preprofile_stage(eng.batch(seqs[:64]), S, gaps)
torch.cuda.synchronize()
res = {"n_seqs": n, "length": L, "tb_gib": os.environ.get("PGPU_TB_GIB", "16"), "streams": os.environ.get("PGPU_DUAL_STREAMS", "2"),
       "tile": os.environ.get("PGPU_TILE_DUAL", "auto")}
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    l0 = eng.launches
    cnt, where, cells = eng.preprofile_stage(batch, S, gaps)
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    res["rep%d" % rep] = {"wall_s": wall, "host_s": t_host, "device_ms": a.elapsed_time(b), "launches": eng.launches - l0,
                          "gcups_reference_cells": cells / wall / 1e9, "gcups_filled": cells / 2 / wall / 1e9}
res["counts_sum"] = int(cnt.sum(dtype=torch.int64).item())
if check:
    sample = np.unique(np.concatenate([np.arange(0, n, max(1, n // 7))[:7], [n - 1]]))
    t0 = time.perf_counter()
    ref, rwhere, rcells = eng.preprofile_stage(batch, S, gaps, masters=sample, shard=None)
    torch.cuda.synchronize()
    res["per_master_path_sample_s"] = time.perf_counter() - t0
    res["per_master_path_gcups"] = rcells / res["per_master_path_sample_s"] / 1e9
    same = True
    for m in sample:
        o1, l1 = where[int(m)]
        o2, l2 = rwhere[int(m)]
        same = same and bool(torch.equal(cnt[o1:o1 + l1 * 27], ref[o2:o2 + l2 * 27]))
    res["sample_identical"] = same
if getattr(eng, "dual_trace", None):
    res["trace_ms_stream_fill0_fill1_walk1"] = [[k, round(a, 2), round(b, 2), round(c, 2)] for k, a, b, c in eng.dual_trace[:14]]
print(json.dumps(res))
