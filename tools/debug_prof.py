import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
import oracle
from praline_b200 import get_engine, matrices, synth, _lib
from praline_b200 import engine as E
eng = get_engine(0)
S = matrices.blosum62()
rng = np.random.default_rng(8)
profs = [synth.profile_from_counts(synth.count_profile(300 + k, int(rng.integers(5, 200)), int(rng.integers(1, 9)), 20, 27)) for k in range(4)]
print([p.shape for p in profs])
pb = eng.profile_batch(profs)
for resident in ("two", "one"):
    for (i, j) in [(0, 1), (2, 3), (1, 0)]:
        got = eng.align_profile_pairs(pb, [i], [j], S, [-11.0, -1.0], mode="global", resident=resident)
        m = oracle.build_scores([profs[i]], [profs[j]], [S])
        g1, g2 = oracle.gap_arrays(profs[i].shape[0], profs[j].shape[0], [-11.0, -1.0])
        want, _ = oracle.align_raw("global", m, g1, g2)
        # also check the rows kernel directly
        transposed = resident == "one"
        res, st = (i, j) if transposed else (j, i)
        K = eng.k_for(int(pb.lens[res])); width = 32 * K
        Ls = int(pb.lens[st])
        rowsrc = np.r_[-1, np.arange(pb.offs[st], pb.offs[st] + Ls)].astype(np.int32)
        rowres = np.full(Ls + 1, res, np.int32)
        mw = torch.zeros((Ls + 1) * width, dtype=torch.float32, device=eng.device)
        Sd, a1, a2 = eng.dev(S), eng.dev(rowsrc), eng.dev(rowres)
        _lib.check(eng.lib.pgpu_build_rows(eng.ptr(pb.prof_dev), eng.ptr(pb.offs_dev), 27, eng.ptr(Sd), eng.ptr(a1),
                                           eng.ptr(a2), Ls + 1, width, int(transposed), 0, eng.ptr(mw), eng.stream()))
        mk = mw.cpu().numpy().reshape(Ls + 1, width)[1:, :int(pb.lens[res])]
        ref = m.T if transposed else m
        print(resident, (i, j), "K", K, "score", float(got[0]), want, "rows equal:", np.array_equal(mk, ref), "maxdiff", np.abs(mk - ref).max())
