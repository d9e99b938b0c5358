"""Device time of K1 alone (k_build_scores_cols at config 5 size) and of the fill alone, operands resident."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth, _lib
L = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = get_engine(0)
S = matrices.nucleotide()
fam = synth.family(5, 16, L, n_sym=4)
def prof(members):
    n = min(len(s) for s in members)
    c = np.zeros((n, S.shape[0]), np.int64)
    for s in members:
        c[np.arange(n), s[:n]] += 1
    return synth.profile_from_counts(c)
p1, p2 = prof(fam[:8]), prof(fam[8:])
d1, d2, dS = eng.dev(p1), eng.dev(p2), eng.dev(S)
L1, L2 = p1.shape[0], p2.shape[0]
m = eng.padded_matrix(L1, L2)
arr = ctypes.c_void_p * 1
A = (ctypes.c_int * 1)(15)
ts = []
for rep in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(eng.lib.pgpu_build_scores(1, arr(d1.data_ptr()), arr(d2.data_ptr()), arr(dS.data_ptr()), A, L1, L2, eng.ptr(m),
                                         int(m.stride(0)), eng.stream()))
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("K1 kernel ms:", [round(t, 3) for t in ts], "nnz/row p1 %.2f p2 %.2f" % ((p1 != 0).sum(1).mean(), (p2 != 0).sum(1).mean()))
