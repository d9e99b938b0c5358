"""Profile x profile all-vs-all (the guide-tree stage on preprofile tracks): exact vs tolerance mode."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth
eng = get_engine(0)
S = matrices.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 200
profs = [synth.profile_from_counts(synth.count_profile(1000 + k, L, depth, 20, 27)) for k in range(n)]
pb = eng.profile_batch(profs)
pi, pj = synth.all_pairs(n)
cells = float((pb.lens[pi] * pb.lens[pj]).sum())
res = {}
only = sys.argv[4] if len(sys.argv) > 4 else ""          # "tc": just the tensor-core pass (for ncu)
for name, fast, tc in (("fast_tc", True, True), ("fast_fma", True, False), ("exact", False, True)):
    if only and not name.endswith(only):
        continue
    eng.fast_tc = tc
    eng.align_profile_pairs(pb, pi, pj, S, [-11.0, -1.0], mode="global", fast=fast)     # warm-up at full size
    eng.take_trace()
    eng.trace_on = True
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc = eng.align_profile_pairs(pb, pi, pj, S, [-11.0, -1.0], mode="global", fast=fast)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = sc
    eng.trace_on = False
    tr = eng.take_trace()
    rows_ms = sum(ms for nm, ms in tr if nm.startswith("score rows"))
    fed_ms = sum(ms for nm, ms in tr if nm == "matrix-fed stream")
    print(json.dumps({"mode": name, "device_ms_score_rows": rows_ms, "device_ms_matrix_fed_stream": fed_ms,
                      "gcups_device": cells / ((rows_ms + fed_ms) * 1e-3) / 1e9}))
    print(json.dumps({"mode": name, "n": n, "L": L, "depth": depth, "pairs": len(pi), "cells": cells,
                      "wall_s": dt, "gcups": cells / dt / 1e9}))
for name in ("fast_tc", "fast_fma"):
    if name in res and "exact" in res:
        rel = np.abs(res[name] - res["exact"]) / np.maximum(1.0, np.abs(res["exact"]))
        print(json.dumps({"max_rel_diff_%s_vs_exact" % name: float(rel.max())}))
