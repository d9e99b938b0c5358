import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from praline_b200 import get_engine, synth, matrices
eng = get_engine(0)
seqs = synth.family(3, 10000, 400)
S = matrices.blosum62()
for rep in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    b = eng.batch(seqs)
    torch.cuda.synchronize(); a = time.perf_counter() - t
    t = time.perf_counter(); flat = np.concatenate([np.asarray(s) for s in seqs]); c1 = time.perf_counter() - t
    t = time.perf_counter(); flat2 = np.concatenate(seqs); c2 = time.perf_counter() - t
    t = time.perf_counter(); f8 = flat.astype(np.uint8); c3 = time.perf_counter() - t
    t = time.perf_counter(); p = torch.from_numpy(f8).pin_memory(); c4 = time.perf_counter() - t
    t = time.perf_counter(); d = p.to("cuda", non_blocking=True); torch.cuda.synchronize(); c5 = time.perf_counter() - t
    t = time.perf_counter(); sd = eng.dev(S); torch.cuda.synchronize(); c6 = time.perf_counter() - t
    t = time.perf_counter(); pl = eng.allpairs_plan(b, S, [-11.0, -1.0], "global", (0, 1)); c7 = time.perf_counter() - t
    print("batch %.2f ms | concat(list-comp) %.2f concat %.2f astype %.2f pin %.2f h2d %.2f devS %.2f plan(cached after 1st) %.2f" % tuple(1e3 * x for x in (a, c1, c2, c3, c4, c5, c6, c7)))
