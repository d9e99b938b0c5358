"""Pins the CPU oracle (oracle/praline_oracle.c) to the reference.

(1) golden vectors produced by the reference's PairwiseAligner / RawPairwiseAligner
    components and cext_align_* (tests/golden, made by tests/golden/make_golden.py);
(2) the reference's own compiled extension, oracle/_ref/cext*.so, on fresh random
    inputs (cell-exact o, t and m) -- skipped only if oracle/_ref was never built.
"""
import os

import numpy as np
import pytest

import oracle
from praline_b200 import matrices, synth
from conftest import MODES, GOLDEN


def _S(case, mats):
    return mats["blosum62"] if case["alphabet"] == "aa" else mats["nucleotide"]


def test_builtin_matrices_match_reference(golden_mats):
    assert np.array_equal(matrices.blosum62(), golden_mats["blosum62"])
    assert np.array_equal(matrices.nucleotide(), golden_mats["nucleotide"])


def test_oracle_vs_golden_sequence_cases(golden_seq, golden_mats):
    assert len(golden_seq) >= 80
    for c in golden_seq:
        S = _S(c, golden_mats)
        a, b = np.asarray(c["a"], np.int32), np.asarray(c["b"], np.int32)
        if c["zero_idxs"] is None:
            score, path = oracle.align_seqs(c["mode"], a, b, S, c["gaps"])
        else:
            m = S[a][:, b]
            g1, g2 = oracle.gap_arrays(len(a), len(b), c["gaps"])
            score, path = oracle.align_raw(c["mode"], m, g1, g2, zero_idxs=[tuple(z) for z in c["zero_idxs"]])
        assert score == c["score"], (c["name"], c["mode"], c["gaps"])
        assert path.tolist() == c["path"], (c["name"], c["mode"], c["gaps"])


def test_survey_known_answers(golden_seq):
    """SURVEY.md section 4 table (scores) is what the golden file holds for 'kat'."""
    want = {("global", 2): 2.0, ("local", 2): 17.0, ("semiglobal_both", 2): 15.0, ("semiglobal_one", 2): 15.0,
            ("semiglobal_two", 2): 15.0, ("global", 1): -8.0, ("local", 1): 20.0, ("semiglobal_both", 1): 17.0}
    seen = 0
    for c in golden_seq:
        if c["name"] == "kat" and (c["mode"], len(c["gaps"])) in want:
            assert c["score"] == want[(c["mode"], len(c["gaps"]))]
            seen += 1
        if c["name"] == "kat_masked":
            assert c["score"] == 13.0 and c["path"] == [[7, 3], [8, 4], [9, 5]]
            seen += 1
    assert seen == 9


def test_oracle_vs_golden_cells(golden_cells):
    for c in golden_cells:
        mode = MODES[int(c["mode"])]
        zero = [tuple(i) for i in np.argwhere(c["z"] != 0)]
        o, t, z = oracle.fill(mode, c["m"], c["g1"], c["g2"], zero_idxs=zero)
        assert np.array_equal(o, c["o"]), mode
        assert np.array_equal(t, c["t"]), mode
        score, path = oracle.align_raw(mode, c["m"], c["g1"], c["g2"], zero_idxs=zero)
        assert score == float(c["score"])
        assert np.array_equal(path, c["path"])


def test_oracle_vs_golden_profiles(golden_prof, golden_mats):
    for c in golden_prof:
        A = int(c["alphabet"])
        S = golden_mats["blosum62"] if A == 27 else golden_mats["nucleotide"]
        p1 = synth.profile_from_counts(c["counts1"])
        p2 = synth.profile_from_counts(c["counts2"])
        m = oracle.build_scores([p1], [p2], [S])
        assert np.array_equal(m, c["m"])
        g1, g2 = oracle.gap_arrays(p1.shape[0], p2.shape[0], list(c["gaps"]))
        score, path = oracle.align_raw(MODES[int(c["mode"])], m, g1, g2)
        assert score == float(c["score"])
        assert np.array_equal(path, c["path"])


needs_ref = pytest.mark.skipif(oracle.ref_cext() is None, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("mode", MODES)
def test_fill_cell_exact_vs_reference_extension(mode):
    rng = np.random.default_rng(hash(mode) % 2**31)
    for trial in range(6):
        L1, L2 = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        if trial % 2:
            m = rng.integers(-4, 12, (L1, L2)).astype(np.float32)
            gaps = [[-11.0, -1.0], [-3.0], [-1.0, -1.0]][trial % 3]
        else:
            m = (rng.standard_normal((L1, L2)) * 3).astype(np.float32)
            gaps = [float(-rng.random() * 8 - 0.1), float(-rng.random() * 2 - 0.01)]
        g1, g2 = oracle.gap_arrays(L1, L2, gaps)
        if trial >= 4:  # per-position gap arrays (RawPairwiseAligner accepts them)
            g1 = (g1 * rng.uniform(0.5, 1.5, g1.shape)).astype(np.float32)
            g2 = (g2 * rng.uniform(0.5, 1.5, g2.shape)).astype(np.float32)
        zero = None
        if trial % 3 == 0:
            zero = [(int(rng.integers(1, L1 + 1)), int(rng.integers(1, L2 + 1))) for _ in range(6)]
        o, t, z = oracle.fill(mode, m, g1, g2, zero)
        ro, rt, rz = oracle.ref_fill(mode, m, g1, g2, zero)
        assert np.array_equal(o, ro)
        assert np.array_equal(t, rt)


@needs_ref
def test_build_scores_bit_exact_vs_reference_extension():
    rng = np.random.default_rng(7)
    for trial in range(5):
        A = [27, 15, 27, 4, 27][trial]
        L1, L2 = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        S = rng.integers(-4, 12, (A, A)).astype(np.float32)
        c1 = rng.integers(0, 4, (L1, A)) * (rng.random((L1, A)) < 0.4)
        c2 = rng.integers(0, 9, (L2, A)) * (rng.random((L2, A)) < 0.3)
        c1[np.arange(L1), rng.integers(0, A, L1)] += 1
        c2[np.arange(L2), rng.integers(0, A, L2)] += 1
        p1, p2 = synth.profile_from_counts(c1), synth.profile_from_counts(c2)
        if trial == 2:  # two track sets summed in set order
            S2 = rng.standard_normal((A, A)).astype(np.float32)
            q1, q2 = p1[:, ::-1].copy(), p2[:, ::-1].copy()
            m = oracle.build_scores([p1, q1], [p2, q2], [S, S2])
            r = oracle.ref_build_scores([p1, q1], [p2, q2], [S, S2])
        else:
            m = oracle.build_scores([p1], [p2], [S])
            r = oracle.ref_build_scores([p1], [p2], [S])
        assert np.array_equal(m, r)


def test_cluster_oracle_vs_reference_golden():
    """oracle.cluster_merge_order against merge orders produced by the reference's own
    HierarchicalClusteringAlgorithm (tests/golden/make_cluster_golden.py)."""
    import json
    with open(os.path.join(GOLDEN, "cluster.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 30
    for c in cases:
        got = oracle.cluster_merge_order(np.asarray(c["dist"], np.float32), c["linkage"])
        assert [list(x) for x in got] == c["order"], (c["n"], c["kind"], c["linkage"])


def test_tree_distance_matrix_matches_reference_expression():
    rng = np.random.default_rng(3)
    n = 9
    sc = rng.integers(-300, -10, n * (n - 1) // 2).astype(np.float32)     # all negative: d.max() is the diagonal 0
    dist = oracle.tree_distance_matrix(sc, n)
    assert dist.dtype == np.float32 and (np.diag(dist) == 0).all() and dist.min() == 0
    iu = np.triu_indices(n, k=1)
    assert np.array_equal(dist[iu], -sc) and np.array_equal(dist, dist.T)


def test_local_master_slave_restatement_against_reference_golden():
    """oracle.align_raw(mode='local', zero_idxs=boxes) and the host restatement of compress_path /
    extend_path_local / Alignment.merge / get_frequencies (tests/we_model.py) reproduce what the
    reference's LocalMasterSlaveAligner + ProfileBuilder produced (tests/golden/local_ms.json)."""
    import json
    import os
    import we_model
    from conftest import GOLDEN
    from praline_b200 import matrices
    S = matrices.blosum62()
    with open(os.path.join(GOLDEN, "local_ms.json")) as f:
        cases = json.load(f)
    for case in cases:
        seqs = [np.asarray(s, np.int32) for s in case["seqs"]]
        for i, gm in enumerate(case["masters"]):
            calls = iter(gm["calls"])
            for j in range(len(seqs)):
                if j == i:
                    continue
                for score, path, n_zero in we_model.we_alignments(seqs[i], seqs[j], S, case["gaps"], case["iterations"]):
                    c = next(calls)
                    assert c["slave"] == "s%d" % j and c["n_zero"] == n_zero
                    assert c["score"] == score
                    assert np.array_equal(np.asarray(c["path"]), path)
            counts, path, names = we_model.local_master_counts(seqs, i, S, case["gaps"], case["iterations"],
                                                               case["threshold"], 27)
            assert np.array_equal(counts, np.asarray(gm["counts"]))
            assert np.array_equal(path, np.asarray(gm["path"]).reshape(len(seqs[i]) + 1, -1))
            assert ["s%d" % k for k in names] == gm["items"]
