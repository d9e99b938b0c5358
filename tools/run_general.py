"""Driver for profiling the general path (K1 + K3): a few profile-profile alignments."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth
eng = get_engine(0)
S = matrices.blosum62()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 300
mode = sys.argv[2] if len(sys.argv) > 2 else "global"
p1 = synth.profile_from_counts(synth.count_profile(1, L, 8, 20, 27))
p2 = synth.profile_from_counts(synth.count_profile(2, L + 7, 8, 20, 27))
g1 = np.empty((L, 2), np.float32); g1[:] = (-11, -1)
g2 = np.empty((L + 7, 2), np.float32); g2[:] = (-11, -1)
for rep in range(3):
    t0 = time.perf_counter()
    m = eng.build_scores([p1], [p2], [S])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    r = eng.align_general(mode, m, g1, g2)
    t2 = time.perf_counter()
    print("L=%d %s: build_scores %.3f ms, align_general %.3f ms, score %.3f" % (L, mode, 1e3 * (t1 - t0), 1e3 * (t2 - t1), r["score"]))
