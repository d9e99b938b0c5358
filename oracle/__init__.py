"""ctypes front-end of the parity oracle (oracle/praline_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of praline_oracle.c.  Importable
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs, nowhere else.  Parity is pinned by tests/test_oracle_pinned.py
against the reference's compiled extension (oracle/_ref) and against golden
vectors made by the reference's PairwiseAligner (tests/golden).
"""
import ctypes
import glob
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {"global": 0, "local": 1, "semiglobal_both": 2, "semiglobal_one": 3, "semiglobal_two": 4}

_lib = None
_ref = None


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "praline_oracle.c")
    stale = not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    have_ref = bool(glob.glob(os.path.join(_HERE, "_ref", "cext*.so")))
    if force or stale or (not have_ref and os.path.exists("/root/reference/praline/util/cext.c")):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        fp = ctypes.POINTER(ctypes.c_float)
        L.orc_align_raw.restype = ctypes.c_int
        L.orc_align_seqs.restype = ctypes.c_int
        L.orc_get_path.restype = ctypes.c_int
        L.orc_extend_semiglobal.restype = ctypes.c_int
        L.orc_align_seqs.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                     ctypes.c_float, fp, ctypes.c_void_p]
        L.orc_align_batch.argtypes = [ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 5 + \
            [ctypes.c_int, ctypes.c_float, ctypes.c_float] + [ctypes.c_void_p] * 4
        L.orc_align_batch.restype = None
        _lib = L
    return _lib


def ref_cext():
    """The reference's own compiled extension (oracle/_ref/cext*.so) or None."""
    global _ref
    if _ref is None:
        build()
        hits = glob.glob(os.path.join(_HERE, "_ref", "cext*.so"))
        if not hits:
            return None
        spec = importlib.util.spec_from_file_location("cext", hits[0])
        _ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ref)
    return _ref


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def build_scores(P1s, P2s, Ss):
    """m = sum_sets P1 . S . P2^T in the reference's evaluation order."""
    n = len(P1s)
    P1s = [_c(p, np.float32) for p in P1s]
    P2s = [_c(p, np.float32) for p in P2s]
    Ss = [_c(s, np.float32) for s in Ss]
    L1, L2 = P1s[0].shape[0], P2s[0].shape[0]
    m = np.zeros((L1, L2), np.float32)
    arr = ctypes.c_void_p * n
    A = (ctypes.c_int * n)(*[p.shape[1] for p in P1s])
    lib().orc_build_scores(n, arr(*[p.ctypes.data for p in P1s]), arr(*[p.ctypes.data for p in P2s]),
                           arr(*[s.ctypes.data for s in Ss]), A, L1, L2, _p(m))
    return m


def gap_arrays(L1, L2, gap_series):
    """component/align.py:182-217: constant per-position gap arrays."""
    gs = list(gap_series)
    if len(gs) == 1:
        gs = [gs[0], gs[0]]
    g1 = np.empty((L1, 2), np.float32)
    g2 = np.empty((L2, 2), np.float32)
    g1[:] = gs
    g2[:] = gs
    return g1, g2


def fill(mode, m, g1, g2, zero_idxs=None):
    """Border init + fill; returns (o, t, z) like the arrays RawPairwiseAligner builds."""
    m = _c(m, np.float32)
    g1 = _c(g1, np.float32)
    g2 = _c(g2, np.float32)
    L1, L2 = m.shape
    o = np.zeros((L1 + 1, L2 + 1, 3), np.float32)
    t = np.zeros((L1 + 1, L2 + 1, 3), np.uint8)
    z = np.zeros((L1 + 1, L2 + 1), np.uint8)
    if zero_idxs is not None:
        for idx in zero_idxs:
            z[idx] = 1
    md = MODES[mode]
    lib().orc_init_borders(md, _p(g1), _p(g2), L1, L2, _p(o), _p(t))
    lib().orc_fill(md, _p(m), _p(g1), _p(g2), L1, L2, _p(o), _p(t), _p(z))
    return o, t, z


def align_raw(mode, m, g1, g2, zero_idxs=None, want_matrices=False):
    """RawPairwiseAligner: returns (score, path[K,2] int32[, o, t])."""
    m = _c(m, np.float32)
    g1 = _c(g1, np.float32)
    g2 = _c(g2, np.float32)
    L1, L2 = m.shape
    path = np.zeros((L1 + L2 + 2, 2), np.int32)
    score = ctypes.c_float()
    zi = None
    nz = 0
    if zero_idxs is not None and len(zero_idxs):
        zi = _c(np.asarray(zero_idxs).reshape(-1, 2), np.int32)
        nz = zi.shape[0]
    o = t = None
    if want_matrices:
        o = np.zeros((L1 + 1, L2 + 1, 3), np.float32)
        t = np.zeros((L1 + 1, L2 + 1, 3), np.uint8)
    n = lib().orc_align_raw(MODES[mode], _p(m), _p(g1), _p(g2), L1, L2,
                            _p(zi) if zi is not None else None, nz, ctypes.byref(score), _p(path),
                            _p(o) if o is not None else None, _p(t) if t is not None else None)
    if want_matrices:
        return float(score.value), path[:n].copy(), o, t
    return float(score.value), path[:n].copy()


def align_seqs(mode, a, b, S, gap_series):
    """PairwiseAligner on two index sequences: (score, path)."""
    a = _c(a, np.int32)
    b = _c(b, np.int32)
    S = _c(S, np.float32)
    gs = list(gap_series)
    if len(gs) == 1:
        gs = [gs[0], gs[0]]
    path = np.zeros((len(a) + len(b) + 2, 2), np.int32)
    score = ctypes.c_float()
    n = lib().orc_align_seqs(MODES[mode], _p(a), len(a), _p(b), len(b), _p(S), S.shape[0],
                             float(gs[0]), float(gs[1]), ctypes.byref(score), _p(path))
    return float(score.value), path[:n].copy()


def align_batch(mode, seqs, offs, pi, pj, S, gap_series, want_paths=False):
    """Scalar CPU loop over pairs; returns scores (and list of paths)."""
    seqs = _c(seqs, np.int32)
    offs = _c(offs, np.int64)
    pi = _c(pi, np.int32)
    pj = _c(pj, np.int32)
    S = _c(S, np.float32)
    gs = list(gap_series)
    if len(gs) == 1:
        gs = [gs[0], gs[0]]
    n = len(pi)
    scores = np.zeros(n, np.float32)
    if want_paths:
        lens = (offs[1:] - offs[:-1])
        cap = lens[pi] + lens[pj] + 2
        poffs = np.zeros(n + 1, np.int64)
        np.cumsum(cap, out=poffs[1:])
        paths = np.zeros((int(poffs[-1]), 2), np.int32)
        plen = np.zeros(n, np.int32)
        lib().orc_align_batch(MODES[mode], n, _p(seqs), _p(offs), _p(pi), _p(pj), _p(S), S.shape[0],
                              float(gs[0]), float(gs[1]), _p(scores), _p(paths), _p(poffs), _p(plen))
        return scores, [paths[poffs[k]:poffs[k] + plen[k]].copy() for k in range(n)]
    lib().orc_align_batch(MODES[mode], n, _p(seqs), _p(offs), _p(pi), _p(pj), _p(S), S.shape[0],
                          float(gs[0]), float(gs[1]), _p(scores), None, None, None)
    return scores


# ---- the reference's compiled extension driven directly (oracle/_ref) -------------------

def ref_init_borders(mode, g1, g2, L1, L2):
    """numpy statement of component/align.py:357-385, used to feed oracle/_ref."""
    o = np.zeros((L1 + 1, L2 + 1, 3), np.float32)
    t = np.zeros((L1 + 1, L2 + 1, 3), np.uint8)
    o[:, 0, :] = -np.inf
    o[0, :, :] = -np.inf
    o[0, 0, 0] = 0
    if mode in ("semiglobal_both", "semiglobal_one"):
        o[:, 0, 1] = 0
    else:
        o[0, 0, 1] = g1[0, 0] - g1[0, 1]
        o[1:, 0, 1] = (np.arange(L1) * g1[:, 1]) + g1[0, 0]
        t[1:, 0, 1] = 1 << 5
    if mode in ("semiglobal_both", "semiglobal_two"):
        o[0, :, 2] = 0
    else:
        o[0, 0, 2] = g2[0, 0] - g2[0, 1]
        o[0, 1:, 2] = (np.arange(L2) * g2[:, 1]) + g2[0, 0]
        t[0, 1:, 2] = 1 << 7
    return o, t


def ref_fill(mode, m, g1, g2, zero_idxs=None):
    """o, t, z filled by the reference's cext_align_<mode> (needs oracle/_ref)."""
    cx = ref_cext()
    m = _c(m, np.float32)
    g1 = _c(g1, np.float32)
    g2 = _c(g2, np.float32)
    L1, L2 = m.shape
    o, t = ref_init_borders(mode, g1, g2, L1, L2)
    z = np.zeros((L1 + 1, L2 + 1), np.uint8)
    if zero_idxs is not None:
        for idx in zero_idxs:
            z[idx] = 1
    getattr(cx, "cext_align_" + mode)(m, g1, g2, o, t, z)
    return o, t, z


def ref_build_scores(P1s, P2s, Ss):
    """m from the reference's cext_build_scores (needs oracle/_ref)."""
    cx = ref_cext()
    P1s = [_c(p, np.float32) for p in P1s]
    P2s = [_c(p, np.float32) for p in P2s]
    Ss = [_c(s, np.float32) for s in Ss]

    def nzmat(i):  # component/align.py:449-458
        r = np.full(i.shape, -1, np.intp)
        for n in range(i.shape[0]):
            row = i[n].nonzero()[0]
            r[n, :row.shape[0]] = row
        return r

    m = np.zeros((P1s[0].shape[0], P2s[0].shape[0]), np.float32)
    cx.cext_build_scores(P1s, P2s, [nzmat(p) for p in P1s], [nzmat(p) for p in P2s], Ss, m)
    return m


# ---- guide-tree clustering (reference: praline/util/cluster.py:27-111) ---------------------------
def cluster_merge_order(dist, linkage="average"):
    """Restatement of HierarchicalClusteringAlgorithm.merge_order (cluster.py:27-57): every
    round rebuilds the float64 inter-cluster matrix `a` over the clusters in ascending id
    (dict insertion order, :34), diagonal 2**32 (:12, :40), takes the first minimum in row-major
    order (:49) and merges cluster `two` into `one` (:55-56).  Linkages (:60-111): min, max, or
    float64 mean of the f32 member distances.  Pinned against the reference class itself
    (tests/golden/cluster.json, made by tests/golden/make_cluster_golden.py)."""
    dist = np.asarray(dist)
    fun = {"single": np.min, "complete": np.max, "average": np.mean}[linkage]
    clusters = {i: [i] for i in range(dist.shape[0])}
    out = []
    while len(clusters) > 1:
        ids = list(clusters)
        a = np.full((len(ids), len(ids)), float(2 ** 32), dtype=float)
        for i, ci in enumerate(ids):
            for j, cj in enumerate(ids):
                if i != j:
                    a[i, j] = fun(dist[np.ix_(clusters[ci], clusters[cj])].astype(float))
        i, j = np.unravel_index(a.argmin(), a.shape)
        one, two = ids[i], ids[j]
        clusters[one] = clusters[one] + clusters[two]
        del clusters[two]
        out.append((int(one), int(two)))
    return out


def tree_distance_matrix(scores_condensed, n):
    """GuideTreeBuilder's distance matrix (component/tree.py:92-147): d f32 [n x n] with 0 on the
    diagonal and the pair scores elsewhere, dist = (-d) + d.max() in f32."""
    d = np.zeros((n, n), np.float32)
    iu = np.triu_indices(n, k=1)
    d[iu] = np.asarray(scores_condensed, np.float32)
    d[(iu[1], iu[0])] = d[iu]
    return (-d) + d.max()
