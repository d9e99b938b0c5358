"""Parity of the CUDA path (through the C ABI) with the oracle and the golden vectors.
All tests here need a B200; they are the parity tests proper."""
import numpy as np
import pytest

import oracle
from praline_b200 import matrices, synth
from conftest import MODES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from praline_b200 import get_engine
    return get_engine(0)


def _check_batch(eng, seqs, pi, pj, S, gaps, mode, resident, paths=True):
    batch = eng.batch(seqs)
    flat, offs = synth.pack(seqs)
    if mode == "local":
        paths = False
    scores, got_paths = eng.align_pairs(batch, pi, pj, S, gaps, mode=mode, want_paths=paths, resident=resident)
    if paths:
        want, want_paths = oracle.align_batch(mode, flat, offs, pi, pj, S, gaps, want_paths=True)
    else:
        want = oracle.align_batch(mode, flat, offs, pi, pj, S, gaps)
    assert np.array_equal(scores, want), (mode, gaps, resident)
    if paths:
        for k, (a, b) in enumerate(zip(got_paths, want_paths)):
            assert np.array_equal(a, b), (mode, gaps, resident, k)


def test_golden_sequence_cases(eng, golden_seq, golden_mats):
    """Every golden vector of the reference's PairwiseAligner, one pair per call."""
    for c in golden_seq:
        S = golden_mats["blosum62"] if c["alphabet"] == "aa" else golden_mats["nucleotide"]
        batch = eng.batch([np.asarray(c["a"]), np.asarray(c["b"])])
        integer = all(float(g).is_integer() for g in c["gaps"])
        if c["zero_idxs"] is None and c["mode"] != "local" and integer:
            for resident in ("one", "two"):
                scores, paths = eng.align_pairs(batch, [0], [1], S, c["gaps"], mode=c["mode"], want_paths=True,
                                                resident=resident)
                assert float(scores[0]) == c["score"], (c["name"], c["mode"], resident)
                assert paths[0].tolist() == c["path"], (c["name"], c["mode"], resident)
        if c["zero_idxs"] is None:
            for resident in ("one", "two"):
                scores, _ = eng.align_pairs(batch, [0], [1], S, c["gaps"], mode=c["mode"], resident=resident)
                assert float(scores[0]) == c["score"], (c["name"], c["mode"], resident)
        zero = None if c["zero_idxs"] is None else [tuple(z) for z in c["zero_idxs"]]
        r = eng.align_seq_pair_general(batch, 0, 1, S, c["gaps"], c["mode"], zero_idxs=zero)
        assert r["score"] == c["score"], (c["name"], c["mode"])
        assert r["path"].tolist() == c["path"], (c["name"], c["mode"])


@pytest.mark.parametrize("resident", ["one", "two"])
@pytest.mark.parametrize("mode", MODES)
def test_batches_vs_oracle(eng, mode, resident):
    S = matrices.blosum62()
    rng = np.random.default_rng(5)
    for trial, (n, length, gaps) in enumerate([(24, 37, [-11.0, -1.0]), (16, 150, [-8.0]), (12, 300, [-11.0, -1.0]),
                                               (10, 90, [-1.0, -1.0]), (8, 420, [-4.0, -2.0])]):
        seqs = synth.family(300 + trial, n, length)
        seqs += [rng.integers(0, 20, int(rng.integers(1, 2 * length))).astype(np.int32) for _ in range(4)]
        pi, pj = synth.all_pairs(len(seqs))
        if trial % 2:   # ordered pairs both ways and self pairs
            pi, pj = np.concatenate([pi, pj[:20], np.arange(5)]), np.concatenate([pj, pi[:20], np.arange(5)])
        _check_batch(eng, seqs, pi, pj, S, gaps, mode, resident)


def test_mixed_length_classes(eng):
    """Pairs whose resident sequences fall in different K classes in one call."""
    S = matrices.blosum62()
    seqs = []
    for k, length in enumerate([5, 33, 70, 129, 200, 260, 330, 390, 450, 600, 700, 900]):
        seqs += synth.family(500 + k, 2, length)
    pi, pj = synth.all_pairs(len(seqs))
    for mode in ("global", "semiglobal_both"):
        _check_batch(eng, seqs, pi, pj, S, [-11.0, -1.0], mode, None)


def test_dna_and_asymmetric_matrix(eng):
    S = matrices.nucleotide().copy()
    S[0, 1] = 3.0   # break symmetry: catches a transposed profile lookup
    seqs = synth.family(9, 10, 80, n_sym=4)
    pi, pj = synth.all_pairs(len(seqs))
    for resident in ("one", "two"):
        for mode in ("global", "semiglobal_one", "semiglobal_two"):
            _check_batch(eng, seqs, pi, pj, S, [-5.0, -2.0], mode, resident)


def test_allpairs_condensed(eng):
    S = matrices.blosum62()
    seqs = synth.family(21, 70, 64) + synth.family(22, 9, 140)
    batch = eng.batch(seqs)
    flat, offs = synth.pack(seqs)
    pi, pj = synth.all_pairs(len(seqs))
    for mode in ("global", "semiglobal_both", "local"):
        out, rng, cells = eng.allpairs_scores(batch, eng.dev(S), 27, [-11.0, -1.0], mode=mode)
        want = oracle.align_batch(mode, flat, offs, pi, pj, S, [-11.0, -1.0])
        assert np.array_equal(out.cpu().numpy(), want), mode
        assert cells == int((batch.lens[pi] * batch.lens[pj]).sum())
    # shards tile the condensed vector without gaps or overlap
    full = out.cpu().numpy()
    acc = np.full_like(full, np.nan)
    for r in range(3):
        o, (lo, hi), _ = eng.allpairs_scores(batch, eng.dev(S), 27, [-11.0, -1.0], mode="local", shard=(r, 3))
        acc[lo:hi] = o.cpu().numpy()[lo:hi]
    assert np.array_equal(acc, full)


def test_fill_debug_matches_golden_cells(eng, golden_cells):
    for c in golden_cells:
        mode = MODES[int(c["mode"])]
        o, t = eng.fill_debug(mode, c["m"], c["g1"], c["g2"], c["z"] if c["z"].any() else None)
        assert np.array_equal(o, c["o"]), mode
        assert np.array_equal(t, c["t"]), mode
        zero = [tuple(i) for i in np.argwhere(c["z"] != 0)]
        r = eng.align_general(mode, c["m"], c["g1"], c["g2"], zero_idxs=zero)
        assert r["score"] == float(c["score"])
        assert np.array_equal(r["path"], c["path"])


@pytest.mark.parametrize("mode", MODES)
def test_general_random_vs_oracle(eng, mode):
    rng = np.random.default_rng(11)
    for trial, (L1, L2) in enumerate([(1, 1), (1, 40), (50, 1), (63, 65), (130, 257), (300, 90), (40, 700)]):
        if trial % 2:
            m = rng.integers(-4, 12, (L1, L2)).astype(np.float32)
            g1, g2 = oracle.gap_arrays(L1, L2, [[-11.0, -1.0], [-3.0]][trial % 4 // 2])
        else:
            m = (rng.standard_normal((L1, L2)) * 3).astype(np.float32)
            g1, g2 = oracle.gap_arrays(L1, L2, [-4.25, -0.625])
            g1 = (g1 * rng.uniform(0.5, 1.5, g1.shape)).astype(np.float32)   # per-position gaps
            g2 = (g2 * rng.uniform(0.5, 1.5, g2.shape)).astype(np.float32)
        zero = None
        if trial in (3, 4):
            zero = [(int(rng.integers(1, L1 + 1)), int(rng.integers(1, L2 + 1))) for _ in range(30)]
        r = eng.align_general(mode, m, g1, g2, zero_idxs=zero, want_matrices=True)
        o, t, z = oracle.fill(mode, m, g1, g2, zero)
        assert np.array_equal(r["o"], o), (mode, trial)
        assert np.array_equal(r["t"], t), (mode, trial)
        ws, wp = oracle.align_raw(mode, m, g1, g2, zero_idxs=zero)
        assert r["score"] == ws, (mode, trial)
        assert np.array_equal(r["path"], wp), (mode, trial)


def test_profiles_vs_golden(eng, golden_prof, golden_mats):
    for c in golden_prof:
        A = int(c["alphabet"])
        S = golden_mats["blosum62"] if A == 27 else golden_mats["nucleotide"]
        p1 = synth.profile_from_counts(c["counts1"])
        p2 = synth.profile_from_counts(c["counts2"])
        m = eng.build_scores([p1], [p2], [S])
        assert np.array_equal(m.cpu().numpy(), c["m"])      # bit-exact, reference evaluation order
        g1, g2 = oracle.gap_arrays(p1.shape[0], p2.shape[0], list(c["gaps"]))
        r = eng.align_general(MODES[int(c["mode"])], m, g1, g2)
        assert abs(r["score"] - float(c["score"])) <= 1e-5 * max(1.0, abs(float(c["score"])))  # stated tolerance
        assert r["score"] == float(c["score"])                                               # and in fact exact
        assert np.array_equal(r["path"], c["path"])


@pytest.mark.parametrize("resident", ["one", "two"])
def test_profile_batches_vs_oracle(eng, resident):
    """Matrix-fed batch: scores of profile x profile pairs are bit-identical to the oracle
    (same evaluation order for m, same recurrence), all modes, ragged lengths."""
    S = matrices.blosum62()
    rng = np.random.default_rng(8)
    profs = []
    for k in range(14):
        L = int(rng.integers(5, 200))
        profs.append(synth.profile_from_counts(synth.count_profile(300 + k, L, int(rng.integers(1, 9)), 20, 27)))
    pb = eng.profile_batch(profs)
    pi, pj = synth.all_pairs(len(profs))
    pi, pj = np.concatenate([pi, pj[:15]]), np.concatenate([pj, pi[:15]])
    for mode, gaps in (("global", [-11.0, -1.0]), ("semiglobal_both", [-2.5, -0.5]), ("local", [-11.0, -1.0]),
                       ("semiglobal_one", [-4.0]), ("semiglobal_two", [-11.0, -1.0])):
        got = eng.align_profile_pairs(pb, pi, pj, S, gaps, mode=mode, resident=resident)
        for k in range(len(pi)):
            p1, p2 = profs[pi[k]], profs[pj[k]]
            m = oracle.build_scores([p1], [p2], [S])
            g1, g2 = oracle.gap_arrays(p1.shape[0], p2.shape[0], gaps)
            want, _ = oracle.align_raw(mode, m, g1, g2)
            assert abs(float(got[k]) - want) <= 1e-5 * max(1.0, abs(want)), (mode, k)   # stated tolerance
            assert float(got[k]) == want, (mode, k)                                     # and in fact exact


def test_preprofile_counts_on_device(eng):
    """Count tables of global master-slave preprofiles from the device equal the reference's
    host pipeline compress_path -> Alignment.merge -> get_frequencies (restated literally here
    from util/align.py:187-232 on oracle paths), with and without a score threshold."""
    S = matrices.blosum62()
    seqs = synth.family(91, 9, 70) + [np.random.default_rng(2).integers(0, 20, 33).astype(np.int32)]
    n = len(seqs)
    batch = eng.batch(seqs)
    masters = np.repeat(np.arange(n), n - 1)
    slaves = np.concatenate([[j for j in range(n) if j != i] for i in range(n)])
    for thr in (None, 60.0):
        cnt, where, scores = eng.preprofile_counts(batch, masters, slaves, S, [-11.0, -1.0], threshold=thr)
        for i in range(n):
            want = np.zeros((len(seqs[i]), 27), np.int64)
            want[np.arange(len(seqs[i])), seqs[i]] += 1
            for j in range(n):
                if j == i:
                    continue
                score, path = oracle.align_seqs("global", seqs[i], seqs[j], S, [-11.0, -1.0])
                if thr is not None and score < thr:
                    continue
                keep = [0] + [r for r in range(1, len(path)) if path[r, 0] > path[r - 1, 0]]   # compress_path
                cp = path[keep]
                for c in range(len(cp) - 1):                                                 # get_frequencies
                    if cp[c + 1, 1] > cp[c, 1]:
                        want[c, seqs[j][cp[c + 1, 1] - 1]] += 1
            off, length = where[i]
            got = cnt[off:off + length * 27].reshape(length, 27)
            assert np.array_equal(got, want), (thr, i)
        pairs_scores = oracle.align_batch("global", *synth.pack(seqs), masters, slaves, S, [-11.0, -1.0])
        assert np.array_equal(scores, pairs_scores)


def test_packed_int16_kernel_matches_f32_kernel(eng):
    """The DPX s16x2 kernel (global, score only, integer scores) and the f32 kernel agree on
    every pair, ragged lengths, both orientations; out-of-range batches fall back to f32."""
    S = matrices.blosum62()
    rng = np.random.default_rng(17)
    seqs = synth.family(401, 40, 180) + [rng.integers(0, 20, int(rng.integers(1, 330))).astype(np.int32) for _ in range(25)]
    batch = eng.batch(seqs)
    pi, pj = synth.all_pairs(len(seqs))
    flat, offs = synth.pack(seqs)
    want = oracle.align_batch("global", flat, offs, pi, pj, S, [-11.0, -1.0])
    for resident in ("one", "two"):
        for use in (True, False):
            eng.use_s16 = use
            got, _ = eng.align_pairs(batch, pi, pj, S, [-11.0, -1.0], mode="global", resident=resident)
            assert np.array_equal(got, want), (resident, use)
    eng.use_s16 = True
    assert eng.fits_s16(S, -11.0, -1.0, batch.lens) is not None
    assert eng.fits_s16(S * 0.5, -11.0, -1.0, batch.lens) is None            # non-integer scores
    assert eng.fits_s16(S * 40, -11.0, -1.0, batch.lens) is None             # would leave int16
    big, _ = eng.align_pairs(batch, pi[:50], pj[:50], S * 40, [-11.0, -1.0], mode="global")
    assert np.array_equal(big, oracle.align_batch("global", flat, offs, pi[:50], pj[:50], S * 40, [-11.0, -1.0]))
    out, _, _ = eng.allpairs_scores(batch, eng.dev(S), 27, [-8.0], mode="global", S_host=S)        # paired residents
    assert np.array_equal(out.cpu().numpy(), oracle.align_batch("global", flat, offs, pi, pj, S, [-8.0]))
    for n_odd in (2, 3, 7, 64):                                    # odd counts, sharded plans
        sub = eng.batch(seqs[:n_odd])
        f2, o2 = synth.pack(seqs[:n_odd])
        qi, qj = synth.all_pairs(n_odd)
        want2 = oracle.align_batch("global", f2, o2, qi, qj, S, [-11.0, -1.0])
        acc = np.full(len(qi), np.nan, np.float32)
        for r in range(2):
            plan = eng.allpairs_tiles(sub, (r, 2), tile=16, paired=True)
            o, (lo, hi), _ = eng.allpairs_scores(sub, eng.dev(S), 27, [-11.0, -1.0], mode="global", shard=(r, 2),
                                                 plan=plan, S_host=S)
            acc[lo:hi] = o.cpu().numpy()[lo:hi]
        assert np.array_equal(acc, want2), n_odd


def test_paired_resident_kernel_every_k_class(eng):
    """The all-vs-all kernel of BASELINE config 3 (k_stream16r<13>: 400-aa residents, two residents
    per register) and every wider instantiation (K = 14 ... 32) against the oracle
    (reference: praline/util/cext.c:99-306 through oracle.align_batch), and against the f32 kernel.
    Families sit just under a K class boundary, with ragged members that fall into lower classes,
    so resident pairs of mixed classes, b_skip tiles and the 8-step blocks with and without
    sequence ends are all exercised; sharded plans must assemble to the same vector."""
    S = matrices.blosum62()
    gaps = [-11.0, -1.0]
    rng = np.random.default_rng(2026)
    cases = [(13, synth.family(3, 44, 400))]                      # config 3's own family (seed 3, 400 aa)
    for K in (13, 14, 16, 20, 24, 32):
        top = 32 * K
        fam = synth.family(500 + K, 7, top - 9)                  # 3 indels of <= 5: stays inside the class
        fam = [s[:top] for s in fam]
        fam += [rng.integers(0, 20, int(n)).astype(np.int32) for n in (top, top - 31, 1, 33, int(rng.integers(2, top)))]
        cases.append((K, fam))
    for K, seqs in cases:
        batch = eng.batch(seqs)
        assert eng.k_for(int(batch.lens.max())) == K
        assert eng.wants_paired(S, gaps[0], gaps[1], 0, batch)
        pi, pj = synth.all_pairs(len(seqs))
        flat, offs = synth.pack(seqs)
        want = oracle.align_batch("global", flat, offs, pi, pj, S, gaps)
        out, _, cells = eng.allpairs_scores(batch, eng.dev(S), 27, gaps, mode="global", S_host=S)
        assert cells == int((batch.lens[pi] * batch.lens[pj]).sum())
        assert np.array_equal(out.cpu().numpy(), want), K
        out32, _, _ = eng.allpairs_scores(batch, eng.dev(S), 27, gaps, mode="global")       # f32 kernel
        assert np.array_equal(out32.cpu().numpy(), want), K
        acc = np.full(len(pi), np.nan, np.float32)
        for r in range(3):                                       # three shards, small tiles
            plan = eng.allpairs_tiles(batch, (r, 3), tile=8, paired=True)
            o, (lo, hi), _ = eng.allpairs_scores(batch, eng.dev(S), 27, gaps, mode="global", shard=(r, 3), plan=plan,
                                                 S_host=S)
            acc[lo:hi] = o.cpu().numpy()[lo:hi]
        assert np.array_equal(acc, want), K


def test_sharded_condensed_layout_and_tree_distance(eng):
    """parallel.ShardedCondensed: every shard writes straight into its slice of the all-gather
    buffer (out_shift); the assembled slices equal the plain condensed vector, and the
    distance-matrix kernel reads the sliced layout (reference: component/tree.py:132-147)."""
    from praline_b200 import parallel
    S = matrices.blosum62()
    gaps = [-11.0, -1.0]
    seqs = synth.family(77, 37, 150)
    n = len(seqs)
    batch = eng.batch(seqs)
    pi, pj = synth.all_pairs(n)
    flat, offs = synth.pack(seqs)
    want = oracle.align_batch("global", flat, offs, pi, pj, S, gaps)
    for world in (1, 2, 4):
        for s_host in (S, None):                                  # packed paired kernel / f32 kernel
            plans = [eng.allpairs_plan(batch, s_host, gaps, "global", (r, world)) for r in range(world)]
            sc = parallel.ShardedCondensed(plans[0][3], eng.device)
            sc.buf.fill_(float("nan"))
            for r in range(world):                               # what each rank does; the all-gather only moves slices
                eng.allpairs_scores(batch, eng.dev(S), 27, gaps, mode="global", shard=(r, world), out=sc,
                                    plan=plans[r], S_host=s_host)
            assert np.array_equal(sc.condensed().cpu().numpy(), want), (world, s_host is None)
            slots = np.arange(len(pi))
            assert np.array_equal(sc.buf.cpu().numpy()[sc.where(slots)], want)
            dist = eng.tree_distance(sc, n)
            assert np.array_equal(dist.cpu().numpy(), oracle.tree_distance_matrix(want, n)), world
    # plain vectors, sizes around the 32 x 32 tiles of the fill kernel, negative-only scores (d.max() is the diagonal's 0)
    for n2 in (1, 2, 31, 32, 33, 65):
        v = np.random.default_rng(n2).integers(-50, 50, n2 * (n2 - 1) // 2).astype(np.float32)
        for vec in (v, -np.abs(v) - 1):
            got = eng.tree_distance(vec, n2).cpu().numpy()
            assert np.array_equal(got, oracle.tree_distance_matrix(vec, n2)), n2


def test_fast_profile_batches_within_tolerance(eng):
    """Tolerance mode of the profile batch (W = P.S^T, A FMAs per cell): every score within the
    stated 1e-5 relative of the oracle, all modes, both orientations."""
    S = matrices.blosum62()
    rng = np.random.default_rng(9)
    profs = [synth.profile_from_counts(synth.count_profile(700 + k, int(rng.integers(20, 260)), int(rng.integers(2, 30)), 20, 27))
             for k in range(12)]
    pb = eng.profile_batch(profs)
    pi, pj = synth.all_pairs(len(profs))
    for resident in ("one", "two"):
        for mode, gaps in (("global", [-11.0, -1.0]), ("semiglobal_both", [-2.5, -0.5]), ("local", [-11.0, -1.0])):
            got = eng.align_profile_pairs(pb, pi, pj, S, gaps, mode=mode, resident=resident, fast=True)
            exact = eng.align_profile_pairs(pb, pi, pj, S, gaps, mode=mode, resident=resident)
            for k in range(len(pi)):
                m = oracle.build_scores([profs[pi[k]]], [profs[pj[k]]], [S])
                g1, g2 = oracle.gap_arrays(m.shape[0], m.shape[1], gaps)
                want, _ = oracle.align_raw(mode, m, g1, g2)
                assert float(exact[k]) == want
                assert abs(float(got[k]) - want) <= 1e-5 * max(1.0, abs(want)), (mode, resident, k, float(got[k]), want)


def test_two_track_sets(eng):
    rng = np.random.default_rng(3)
    S1, S2 = matrices.blosum62(), rng.standard_normal((15, 15)).astype(np.float32)
    p1 = synth.profile_from_counts(synth.count_profile(1, 33, 5, 20, 27))
    p2 = synth.profile_from_counts(synth.count_profile(2, 47, 7, 20, 27))
    q1 = synth.profile_from_counts(synth.count_profile(3, 33, 3, 4, 15))
    q2 = synth.profile_from_counts(synth.count_profile(4, 47, 2, 4, 15))
    m = eng.build_scores([p1, q1], [p2, q2], [S1, S2]).cpu().numpy()
    assert np.array_equal(m, oracle.build_scores([p1, q1], [p2, q2], [S1, S2]))


def test_full_size_properties(eng):
    """BASELINE config 2 shape (1,000 x 300 aa): properties that need no oracle pass, plus a
    sample of pairs checked against the oracle."""
    S = matrices.blosum62()
    seqs = synth.family(2, 1000, 300)
    batch = eng.batch(seqs)
    n = len(seqs)
    out, _, cells = eng.allpairs_scores(batch, eng.dev(S), 27, [-11.0, -1.0], mode="global", S_host=S)   # packed kernel
    full = out.cpu().numpy()
    out32, _, _ = eng.allpairs_scores(batch, eng.dev(S), 27, [-11.0, -1.0], mode="global")                # f32 kernel
    assert np.array_equal(out32.cpu().numpy(), full)
    pi, pj = synth.all_pairs(n)
    assert cells == int((batch.lens[pi] * batch.lens[pj]).sum())
    # (1) the other orientation (resident = sequence two) gives the same scores
    rng = np.random.default_rng(0)
    pick = rng.choice(len(pi), 4000, replace=False)
    s2, _ = eng.align_pairs(batch, pi[pick], pj[pick], S, [-11.0, -1.0], mode="global", resident="two")
    assert np.array_equal(s2, full[pick])
    # (2) global alignment score is symmetric for a symmetric matrix
    s3, _ = eng.align_pairs(batch, pj[pick], pi[pick], S, [-11.0, -1.0], mode="global")
    assert np.array_equal(s3, full[pick])
    # (3) a sequence against itself scores the sum of its diagonal substitution scores
    ii = np.arange(0, n, 7)
    s4, p4 = eng.align_pairs(batch, ii, ii, S, [-11.0, -1.0], mode="global", want_paths=True)
    assert np.array_equal(s4, np.asarray([S[seqs[i], seqs[i]].sum() for i in ii], np.float32))
    assert all(np.array_equal(p, np.stack([np.arange(len(seqs[i]) + 1)] * 2, 1)) for p, i in zip(p4, ii))
    # (4) oracle on a sample, with paths
    flat, offs = synth.pack(seqs)
    sm = pick[:300]
    want, wpaths = oracle.align_batch("global", flat, offs, pi[sm], pj[sm], S, [-11.0, -1.0], want_paths=True)
    assert np.array_equal(full[sm], want)
    got, gpaths = eng.align_pairs(batch, pi[sm], pj[sm], S, [-11.0, -1.0], mode="global", want_paths=True)
    assert np.array_equal(got, want)
    assert all(np.array_equal(a, b) for a, b in zip(gpaths, wpaths))
    # (5) a traced path re-scores to the reported score
    for k in range(0, 300, 37):
        a, b, p = seqs[pi[sm[k]]], seqs[pj[sm[k]]], gpaths[k]
        sc, gap = 0.0, None
        for (y0, x0), (y1, x1) in zip(p[:-1], p[1:]):
            if y1 > y0 and x1 > x0:
                sc += S[a[y1 - 1], b[x1 - 1]]; gap = None
            else:
                kind = "u" if y1 > y0 else "l"
                sc += -1.0 if gap == kind else -11.0
                gap = kind
        assert sc == got[k]


def test_long_profile_alignment_vs_oracle(eng):
    """BASELINE config 5 shape, scaled to what the oracle finishes in seconds: two depth-8 DNA
    count profiles (nucleotide matrix), global + semiglobal_both, affine [-11,-1] and linear [-2]."""
    S = matrices.nucleotide()
    c1 = synth.count_profile(5, 2600, 8, 4, 15)
    c2 = synth.count_profile(6, 3100, 8, 4, 15)
    p1, p2 = synth.profile_from_counts(c1), synth.profile_from_counts(c2)
    m = eng.build_scores([p1], [p2], [S])
    want_m = oracle.build_scores([p1], [p2], [S])
    assert np.array_equal(m.cpu().numpy(), want_m)
    for mode in ("global", "semiglobal_both"):
        for gaps in ([-11.0, -1.0], [-2.0]):
            g1, g2 = oracle.gap_arrays(2600, 3100, gaps)
            r = eng.align_general(mode, m, g1, g2)
            ws, wp = oracle.align_raw(mode, want_m, g1, g2)
            assert r["score"] == ws, (mode, gaps)
            assert np.array_equal(r["path"], wp), (mode, gaps)


def test_profile_pair_in_one_call_vs_oracle(eng, monkeypatch):
    """pgpu_align_profile_long (Engine.align_profile_pair): K1 beside the wavefront fill -- per-block ready flags
    instead of a finished matrix -- gives the oracle's score and path on the config 5 shape (2600 x 3100: 21 x 25
    blocks of m, the last ones ragged), the sequential path the same, and small or non-eligible inputs (local mode,
    varying gaps, matrices below 2^21 cells) take the sequential path inside the same call."""
    S = matrices.nucleotide()
    p1 = synth.profile_from_counts(synth.count_profile(5, 2600, 8, 4, 15))
    p2 = synth.profile_from_counts(synth.count_profile(6, 3100, 8, 4, 15))
    want_m = oracle.build_scores([p1], [p2], [S])
    for mode in ("global", "semiglobal_both"):
        for gaps in ([-11.0, -1.0], [-2.0]):
            g1, g2 = oracle.gap_arrays(2600, 3100, gaps)
            ws, wp = oracle.align_raw(mode, want_m, g1, g2)
            for rep in range(2):                    # the flags are cleared per call
                r = eng.align_profile_pair(mode, p1, p2, S, g1, g2)
                assert r["score"] == ws and np.array_equal(r["path"], wp), (mode, gaps, rep)
            monkeypatch.setenv("PGPU_NO_K1_OVERLAP", "1")
            r = eng.align_profile_pair(mode, p1, p2, S, g1, g2)
            monkeypatch.delenv("PGPU_NO_K1_OVERLAP")
            assert r["score"] == ws and np.array_equal(r["path"], wp), (mode, gaps)
    rng = np.random.default_rng(2)
    g1, g2 = oracle.gap_arrays(2600, 3100, [-11.0, -1.0])
    g1v = g1.copy()
    g1v[7] = (-3.0, -0.5)                           # per-position gaps: k_wave, K1 first
    for mode, ga in (("local", g1), ("global", g1v)):
        r = eng.align_profile_pair(mode, p1, p2, S, ga, g2)
        ws, wp = oracle.align_raw(mode, want_m, ga, g2)
        assert r["score"] == ws and np.array_equal(r["path"], wp), mode
    q1 = synth.profile_from_counts(synth.count_profile(15, int(rng.integers(40, 90)), 5, 20, 27))
    q2 = synth.profile_from_counts(synth.count_profile(16, int(rng.integers(40, 90)), 5, 20, 27))
    B = matrices.blosum62()
    h1, h2 = oracle.gap_arrays(q1.shape[0], q2.shape[0], [-11.0, -1.0])
    r = eng.align_profile_pair("semiglobal_both", q1, q2, B, h1, h2)
    ws, wp = oracle.align_raw("semiglobal_both", oracle.build_scores([q1], [q2], [B]), h1, h2)
    assert r["score"] == ws and np.array_equal(r["path"], wp)


def test_large_score_matrix_kernel_edge_shapes(eng):
    """k_build_scores_cols (column-per-thread K1 for matrices of 2M cells and more) is bit-identical to the
    oracle's evaluation order (cext.c:63-95, :388-421) on ragged shapes, sparse DNA profiles (table path) and dense
    amino-acid profiles whose columns need more quads than the table holds (direct path)."""
    for seed, (L1, L2, depth, nsym, A, S) in enumerate([(1500, 1411, 8, 4, 15, matrices.nucleotide()),
                                                       (129, 16400, 3, 4, 15, matrices.nucleotide()),
                                                       (1900, 1153, 40, 20, 27, matrices.blosum62())]):
        p1 = synth.profile_from_counts(synth.count_profile(50 + seed, L1, depth, nsym, A))
        p2 = synth.profile_from_counts(synth.count_profile(60 + seed, L2, depth, nsym, A))
        m = eng.build_scores([p1], [p2], [S])
        assert np.array_equal(m.cpu().numpy(), oracle.build_scores([p1], [p2], [S])), (L1, L2, A)


@pytest.mark.parametrize("mode", ["global", "semiglobal_both", "semiglobal_one", "semiglobal_two"])
def test_row_blocked_wavefront_vs_oracle(eng, mode):
    """k_wave4 (4 rows x 4 columns per lane and step) against the oracle: row counts around the block size,
    several strips, a last strip with idle lanes, tie-heavy integer scores and f32 profile scores."""
    rng = np.random.default_rng(3)
    S = matrices.nucleotide()
    for L1, L2 in [(1, 1), (2, 5), (3, 129), (4, 128), (5, 127), (7, 300), (257, 515), (1001, 777)]:
        a_, b_ = rng.integers(0, 4, L1), rng.integers(0, 4, L2)
        m = np.ascontiguousarray(S[a_][:, b_])
        if L1 > 200:
            m = m + rng.integers(-2, 3, m.shape).astype(np.float32) * 0.25       # non-integer, still tie-prone
        for gaps in ([-11.0, -1.0], [-2.0], [-0.5, -0.5]):
            g1, g2 = oracle.gap_arrays(L1, L2, gaps)
            r = eng.align_general(mode, m, g1, g2)
            ws, wp = oracle.align_raw(mode, m, g1, g2)
            assert r["score"] == ws, (mode, L1, L2, gaps)
            assert np.array_equal(r["path"], wp), (mode, L1, L2, gaps)


def test_config5_full_size_properties(eng):
    """20,000 x 20,000 (BASELINE config 5): too large for the oracle in a test, so size-independent
    properties: the traced path is monotone, spans the matrix and re-scores to the reported score;
    a linear-gap score is bounded by the affine one with the same extension."""
    S = matrices.nucleotide()
    L = 20000
    a = synth.family(5, 2, L, n_sym=4, n_indels=0)
    seqs = [a[0], a[1]]
    batch = eng.batch(seqs)
    scores = {}
    for gaps in ([-11.0, -1.0], [-1.0]):
        r = eng.align_seq_pair_general(batch, 0, 1, S, gaps, "global")
        p = r["path"]
        assert tuple(p[0]) == (0, 0) and tuple(p[-1]) == (L, L)
        d = np.diff(p, axis=0)
        assert ((d >= 0).all() and (d.max(axis=1) == 1).all())
        go, ge = (gaps + gaps)[:2]
        sc, gap = 0.0, None
        diag = (d[:, 0] == 1) & (d[:, 1] == 1)
        sc = float(S[seqs[0][p[1:, 0][diag] - 1], seqs[1][p[1:, 1][diag] - 1]].sum())
        kinds = np.where(diag, 0, np.where(d[:, 0] == 1, 1, 2))
        prev = np.r_[0, kinds[:-1]]
        opens = (kinds != 0) & (kinds != prev)
        exts = (kinds != 0) & (kinds == prev)
        sc += go * opens.sum() + ge * exts.sum()
        assert sc == r["score"], gaps
        scores[tuple(gaps)] = r["score"]
    assert scores[(-1.0,)] >= scores[(-11.0, -1.0)]


def test_length_boundaries(eng):
    """Resident lengths on both sides of every columns-per-lane class boundary (32*K), traced and
    score-only, f32 and packed kernels; plus residents beyond the inter-task limit through the
    general kernels."""
    S = matrices.blosum62()
    rng = np.random.default_rng(23)
    lens = [1, 2, 31, 32, 33, 63, 64, 65, 96, 97, 128, 129, 192, 193, 256, 257, 319, 320, 321, 384, 385,
            416, 417, 448, 449, 512, 513, 640, 641, 768, 769, 1023, 1024]
    seqs = [rng.integers(0, 20, n).astype(np.int32) for n in lens]
    partner = synth.family(7, 3, 150)
    seqs += partner
    n0 = len(lens)
    pi = np.concatenate([np.arange(n0), np.full(n0, n0), np.arange(n0 - 1)])
    pj = np.concatenate([np.full(n0, n0 + 1), np.arange(n0), np.arange(1, n0)])
    flat, offs = synth.pack(seqs)
    batch = eng.batch(seqs)
    for mode in ("global", "semiglobal_both"):
        want, wpaths = oracle.align_batch(mode, flat, offs, pi, pj, S, [-11.0, -1.0], want_paths=True)
        for resident in ("one", "two"):
            for use in (True, False):
                eng.use_s16 = use
                got, _ = eng.align_pairs(batch, pi, pj, S, [-11.0, -1.0], mode=mode, resident=resident)
                assert np.array_equal(got, want), (mode, resident, use)
            eng.use_s16 = True
            got, paths = eng.align_pairs(batch, pi, pj, S, [-11.0, -1.0], mode=mode, want_paths=True, resident=resident)
            assert np.array_equal(got, want)
            assert all(np.array_equal(a, b) for a, b in zip(paths, wpaths)), (mode, resident)
    # beyond 1024 columns: the inter-task kernel refuses, the general path serves
    long_seqs = [rng.integers(0, 20, 1500).astype(np.int32), rng.integers(0, 20, 1100).astype(np.int32)]
    lb = eng.batch(long_seqs)
    with pytest.raises(Exception):
        eng.align_pairs(lb, [0], [1], S, [-11.0, -1.0], mode="global")
    r = eng.align_seq_pair_general(lb, 0, 1, S, [-11.0, -1.0], "global")
    ws, wp = oracle.align_seqs("global", long_seqs[0], long_seqs[1], S, [-11.0, -1.0])
    assert r["score"] == ws and np.array_equal(r["path"], wp)


# ---- guide-tree clustering (csrc/cluster.cu vs util/cluster.py) ---------------------------------
def test_cluster_kernel_vs_reference_golden(eng):
    import json, os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "cluster.json")) as f:
        cases = json.load(f)
    for c in cases:
        got = eng.cluster_merge_order(np.asarray(c["dist"], np.float32), c["linkage"])
        assert [list(x) for x in got] == c["order"], (c["n"], c["kind"], c["linkage"])


@pytest.mark.parametrize("linkage", ["single", "complete", "average"])
def test_cluster_kernel_vs_oracle_ties(eng, linkage):
    """Tie-heavy integer distances (what sequence scores give): the order is exact."""
    rng = np.random.default_rng(41)
    for n, hi in ((37, 4), (70, 12)):
        s = np.triu(rng.integers(0, hi, (n, n)), 1).astype(np.float32)
        dist = s + s.T
        assert eng.cluster_merge_order(dist, linkage) == oracle.cluster_merge_order(dist, linkage)


def _cluster_numpy_fast(dist, linkage):
    """Independent O(n^3 / vectorised) check for large n: float64 sum / min / max matrix updated in
    place, argmin over the live sub-matrix in ascending-id order."""
    n = dist.shape[0]
    S = dist.astype(np.float64)
    cnt = np.ones(n)
    ids = list(range(n))
    out = []
    while len(ids) > 1:
        sub = S[np.ix_(ids, ids)]
        a = sub / np.outer(cnt[ids], cnt[ids]) if linkage == "average" else sub.copy()
        np.fill_diagonal(a, float(2 ** 32))
        i, j = np.unravel_index(a.argmin(), a.shape)
        one, two = ids[i], ids[j]
        if linkage == "average":
            S[one, :] += S[two, :]
        elif linkage == "single":
            S[one, :] = np.minimum(S[one, :], S[two, :])
        else:
            S[one, :] = np.maximum(S[one, :], S[two, :])
        S[:, one] = S[one, :]
        cnt[one] += cnt[two]
        ids.pop(j)
        out.append((one, two))
    return out


def test_cluster_kernel_large_from_real_scores(eng):
    """600 sequences: all-vs-all scores -> distance matrix on the device -> clustering kernel,
    against the vectorised restatement (integer scores: exact) and structural properties."""
    S = matrices.blosum62()
    seqs = synth.family(9, 600, 60)
    n = len(seqs)
    batch = eng.batch(seqs)
    cond, _, _ = eng.allpairs_scores(batch, eng.dev(S), 27, [-11.0, -1.0], mode="global", S_host=S)
    dist = eng.tree_distance(cond, n)
    want_dist = oracle.tree_distance_matrix(cond.cpu().numpy(), n)
    assert np.array_equal(dist.cpu().numpy(), want_dist)
    for linkage in ("average", "single", "complete"):
        got = eng.cluster_merge_order(dist, linkage)
        assert got == _cluster_numpy_fast(want_dist, linkage), linkage
        twos = [b for _, b in got]
        assert len(got) == n - 1 and len(set(twos)) == n - 1       # every cluster is merged away once
        gone = set()
        for a, b in got:
            assert a not in gone and b not in gone and a != b
            gone.add(b)


@pytest.mark.parametrize("resident", ["one", "two"])
def test_packed_traced_kernel_paths_and_counts(eng, resident):
    """Global traced batches with integer scores inside +-16000 run the packed int16 traced kernel
    (tb_fmt 1): paths equal the oracle's and the f32 traced kernel's, for both orientations, mixed
    lengths across K classes, self pairs and both pair orders."""
    S = matrices.blosum62()
    rng = np.random.default_rng(77)
    seqs = synth.family(71, 14, 180) + synth.family(72, 9, 75) + [rng.integers(0, 20, n).astype(np.int32) for n in (1, 2, 33, 64, 250, 400)]
    pi, pj = synth.all_pairs(len(seqs))
    pi, pj = np.concatenate([pi, pj[:60], np.arange(6)]), np.concatenate([pj, pi[:60], np.arange(6)])
    flat, offs = synth.pack(seqs)
    batch = eng.batch(seqs)
    for gaps in ([-11.0, -1.0], [-3.0], [-2.0, -2.0]):
        want, wpaths = oracle.align_batch("global", flat, offs, pi, pj, S, gaps, want_paths=True)
        eng.use_s16 = True
        got, paths = eng.align_pairs(batch, pi, pj, S, gaps, mode="global", want_paths=True, resident=resident)
        assert eng.last_traced_fmt == 1
        assert np.array_equal(got, want)
        assert all(np.array_equal(a, b) for a, b in zip(paths, wpaths)), gaps
        eng.use_s16 = False
        try:
            got32, paths32 = eng.align_pairs(batch, pi, pj, S, gaps, mode="global", want_paths=True, resident=resident)
            assert eng.last_traced_fmt == 0
        finally:
            eng.use_s16 = True
        assert np.array_equal(got32, want) and all(np.array_equal(a, b) for a, b in zip(paths32, wpaths))
    # beyond the +-16000 working range the f32 kernel takes over
    long_seqs = synth.family(73, 4, 900)
    lb = eng.batch(long_seqs)
    eng.align_pairs(lb, [0, 1], [2, 3], S, [-11.0, -1.0], mode="global", want_paths=True, resident=resident)
    assert eng.last_traced_fmt == 0


def _we_family():
    rng = np.random.default_rng(5)
    fam = [np.asarray(s) for s in synth.family(17, 6, 90)] + [rng.integers(0, 20, 15).astype(np.int32)]
    fam += [np.asarray(s) for s in synth.family(19, 3, 150)]
    w = 17   # tryptophan: S[W][W] = 11 ties a gap opening of -11 against the local zero
    fam.append(np.array([w, 3, w, w, 5, w, 2, w, w], np.int32))
    fam.append(np.array([w, w, 4, w, 6, 6, w, w, 1, w], np.int32))
    return fam


@pytest.mark.parametrize("gaps", [[-11.0, -1.0], [-3.0], [-11.0, 0.0]])
def test_local_waterman_eggert_batches_vs_oracle(eng, gaps):
    """Batched local alignments with Waterman-Eggert iterations (pgpu_align_tiles_local +
    pgpu_traceback_tiles_local): score, path and bounding box of every (pair, iteration) equal the
    reference's LocalMasterSlaveAligner inner loop restated on the oracle (tests/we_model.py)."""
    import we_model
    S = matrices.blosum62()
    fam = _we_family()
    n = len(fam)
    batch = eng.batch(fam)
    pi = np.repeat(np.arange(n), n - 1)
    pj = np.concatenate([[j for j in range(n) if j != i] for i in range(n)])
    for iterations in (1, 4):
        scores, paths, boxes = eng.local_pairs(batch, pi, pj, S, gaps, iterations=iterations, want_paths=True)
        for k, (i, j) in enumerate(zip(pi, pj)):
            want = we_model.we_alignments(fam[i], fam[j], S, gaps, iterations)
            for it, (score, path, _) in enumerate(want):
                assert scores[it][k] == score, (gaps, i, j, it)
                assert np.array_equal(paths[it][k], path), (gaps, i, j, it)
            wb = we_model.boxes_of([p for _, p, _ in want])
            for it in range(min(iterations, 3)):
                assert tuple(boxes[k, it]) == wb[it], (gaps, i, j, it)


def test_local_batches_vs_reference_golden(eng):
    """The same against what the reference itself produced (tests/golden/local_ms.json)."""
    import json
    import os
    from conftest import GOLDEN
    S = matrices.blosum62()
    with open(os.path.join(GOLDEN, "local_ms.json")) as f:
        cases = json.load(f)
    for case in cases:
        seqs = [np.asarray(s, np.int32) for s in case["seqs"]]
        n = len(seqs)
        batch = eng.batch(seqs)
        masters = np.repeat(np.arange(n), n - 1)
        slaves = np.concatenate([[j for j in range(n) if j != i] for i in range(n)])
        scores, paths, _ = eng.local_pairs(batch, masters, slaves, S, case["gaps"], iterations=case["iterations"],
                                           want_paths=True)
        cnt, where, sc2 = eng.local_preprofile_counts(batch, masters, slaves, S, case["gaps"],
                                                      iterations=case["iterations"], threshold=case["threshold"])
        assert np.array_equal(scores, sc2)
        k = 0
        for i, gm in enumerate(case["masters"]):
            calls = iter(gm["calls"])
            for j in range(n):
                if j == i:
                    continue
                for it in range(case["iterations"]):
                    c = next(calls)
                    assert c["score"] == scores[it][k]
                    assert np.array_equal(np.asarray(c["path"]), paths[it][k])
                k += 1
            off, length = where[i]
            assert np.array_equal(cnt[off:off + length * 27].reshape(length, 27), np.asarray(gm["counts"])), (case["seed"], i)


@pytest.mark.parametrize("thr", [None, 35.0])
def test_local_preprofile_counts_on_device(eng, thr):
    """Count tables of local master-slave preprofiles from the device equal the reference's host
    pipeline compress_path -> extend_path_local -> Alignment.merge -> get_frequencies."""
    import we_model
    S = matrices.blosum62()
    fam = _we_family()
    n = len(fam)
    batch = eng.batch(fam)
    masters = np.repeat(np.arange(n), n - 1)
    slaves = np.concatenate([[j for j in range(n) if j != i] for i in range(n)])
    cnt, where, scores = eng.local_preprofile_counts(batch, masters, slaves, S, [-11.0, -1.0], iterations=2, threshold=thr)
    for i in range(n):
        want, _, _ = we_model.local_master_counts(fam, i, S, [-11.0, -1.0], 2, thr, 27)
        off, length = where[i]
        assert np.array_equal(cnt[off:off + length * 27].reshape(length, 27), want), (thr, i)


def test_local_batch_full_size_properties(eng):
    """At a size the oracle does not finish quickly: iteration 1 of the local batch scores what the
    score-only local kernel scores, every
    path lies inside its own box and outside the boxes of earlier iterations (first cell excepted:
    a walk may END on a masked cell), and the count tables hold one count per (pair, iteration,
    aligned master position) at most."""
    S = matrices.blosum62()
    seqs = synth.family(23, 120, 300)
    n = len(seqs)
    batch = eng.batch(seqs)
    pi = np.repeat(np.arange(n), n - 1)
    pj = np.concatenate([[j for j in range(n) if j != i] for i in range(n)])
    scores, paths, boxes = eng.local_pairs(batch, pi, pj, S, [-11.0, -1.0], iterations=3, want_paths=True)
    plain, _ = eng.align_pairs(batch, pi, pj, S, [-11.0, -1.0], mode="local")
    assert np.array_equal(scores[0], plain)
    assert (scores >= 0).all()
    rng = np.random.default_rng(0)
    for k in rng.integers(0, len(pi), 400):
        for it in range(3):
            p = paths[it][k]
            ylo, yhi, xlo, xhi = boxes[k, it]
            assert (p[0] == (ylo, xlo)).all() and (p[-1] == (yhi, xhi)).all()
            d = np.diff(p, axis=0)
            assert ((d >= 0) & (d <= 1)).all() and (d.sum(axis=1) >= 1).all()
            for e in range(it):
                b = boxes[k, e]
                inside = (p[1:, 0] >= max(b[0], 1)) & (p[1:, 0] <= b[1]) & (p[1:, 1] >= max(b[2], 1)) & (p[1:, 1] <= b[3])
                assert not inside.any()


@pytest.mark.parametrize("mode", ["global", "local"])
def test_preprofile_stage_chunks_equal_one_batch(eng, mode):
    """Engine.preprofile_stage (masters in chunks, nothing read back between chunks) gives the count
    tables of the one-batch entry points, for both master-slave aligners."""
    S = matrices.blosum62()
    seqs = synth.family(29, 23, 80) + [np.random.default_rng(1).integers(0, 20, 40).astype(np.int32)]
    n = len(seqs)
    batch = eng.batch(seqs)
    masters = np.repeat(np.arange(n), n - 1)
    slaves = np.concatenate([[j for j in range(n) if j != i] for i in range(n)])
    if mode == "global":
        want, where, _ = eng.preprofile_counts(batch, masters, slaves, S, [-11.0, -1.0], threshold=50.0)
    else:
        want, where, _ = eng.local_preprofile_counts(batch, masters, slaves, S, [-11.0, -1.0], iterations=2, threshold=50.0)
    cnt, where2, cells = eng.preprofile_stage(batch, S, [-11.0, -1.0], threshold=50.0, mode=mode, iterations=2,
                                              chunk_pairs=5 * (n - 1))
    assert where == where2
    assert np.array_equal(cnt.cpu().numpy(), want)
    lens = batch.lens
    assert cells == int((lens.sum() ** 2 - (lens ** 2).sum())) * (2 if mode == "local" else 1)


@pytest.mark.parametrize("gaps", [[-11.0, -1.0], [-2.0], [-4.0, -2.0], [0.0, 0.0]])
def test_dual_traced_fill_serves_both_orientations(eng, gaps):
    """Paired-resident traced kernel + dual walk (gotoh_stream16r.cuh, tb_fmt 2): ONE fill of an unordered
    pair must give the reference's path for (i, j) AND for (j, i) -- tie order included (util/align.py:161-174)
    -- on tie-heavy inputs (three-letter sequences, linear and zero gaps), ragged lengths over several
    columns-per-lane classes, odd sequence counts, several waves and both streams."""
    S = matrices.blosum62()
    assert np.array_equal(S, S.T)
    rng = np.random.default_rng(int(-gaps[0]) + 5)
    seqs = synth.family(77, 9, 90)
    seqs += [rng.integers(0, 3, int(rng.integers(1, 140))).astype(np.int32) for _ in range(10)]       # tie-heavy
    seqs += [rng.integers(0, 20, L).astype(np.int32) for L in (1, 2, 33, 64, 65, 129, 193, 260)]
    batch = eng.batch(seqs)
    flat, offs = synth.pack(seqs)
    n = len(seqs)
    budget = eng.tb_budget_words
    eng.tb_budget_words = 1 << 20            # force several waves
    try:
        scores, paths, cells = eng.allpairs_dual(batch, S, gaps, want_paths=True, tile=16)
    finally:
        eng.tb_budget_words = budget
    pi, pj = synth.all_pairs(n)
    assert cells == int((batch.lens[pi] * batch.lens[pj]).sum())
    both_i = np.concatenate([pi, pj])
    both_j = np.concatenate([pj, pi])
    want, want_paths = oracle.align_batch("global", flat, offs, both_i, both_j, S, gaps, want_paths=True)
    assert len(paths) == 2 * len(pi)
    differ = 0
    for k, (i, j) in enumerate(zip(both_i, both_j)):
        i, j = int(i), int(j)
        assert scores[(min(i, j), max(i, j))] == want[k], (i, j)
        assert np.array_equal(paths[(i, j)], want_paths[k]), (gaps, i, j)
    for i, j in zip(pi, pj):
        differ += not np.array_equal(paths[(int(i), int(j))][:, ::-1], paths[(int(j), int(i))])
    assert differ > 0          # the two orientations really break ties differently on these inputs


def test_dual_traced_long_residents(eng):
    """The same on the K classes of BASELINE config 3 and beyond (400-aa residents: K = 13; up to K = 20 -- the
    traced packed range, +-16000, ends near 660 residues with BLOSUM62 and [-11, -1])."""
    S = matrices.blosum62()
    rng = np.random.default_rng(5)
    seqs = synth.family(3, 7, 400) + [rng.integers(0, 20, L).astype(np.int32) for L in (417, 448, 512, 640)]
    batch = eng.batch(seqs)
    flat, offs = synth.pack(seqs)
    assert eng.dual_traced_ok(S, [-11.0, -1.0], batch.lens) is not None
    scores, paths, _ = eng.allpairs_dual(batch, S, [-11.0, -1.0], want_paths=True, tile=16)
    pi, pj = synth.all_pairs(len(seqs))
    both_i, both_j = np.concatenate([pi, pj]), np.concatenate([pj, pi])
    want, want_paths = oracle.align_batch("global", flat, offs, both_i, both_j, S, [-11.0, -1.0], want_paths=True)
    for k, (i, j) in enumerate(zip(both_i, both_j)):
        assert scores[(int(min(i, j)), int(max(i, j)))] == want[k]
        assert np.array_equal(paths[(int(i), int(j))], want_paths[k]), (i, j)


def test_preprofile_stage_symmetric_path_equals_per_master_path(eng):
    """Engine.preprofile_stage: the symmetric path (one fill per unordered pair, two walks) and the per-master
    path (one fill per ordered pair) give identical count tables; sharded over two ranks the tables add up."""
    S = matrices.blosum62()
    rng = np.random.default_rng(12)
    seqs = synth.family(31, 37, 120) + [rng.integers(0, 4, int(rng.integers(5, 90))).astype(np.int32) for _ in range(6)]
    batch = eng.batch(seqs)
    for thr in (None, 40.0):
        want, where, cells = eng.preprofile_stage(batch, S, [-11.0, -1.0], threshold=thr, shard=None)
        got, where2, cells2 = eng.preprofile_stage(batch, S, [-11.0, -1.0], threshold=thr)
        assert where == where2 and cells == cells2
        assert np.array_equal(got.cpu().numpy(), want.cpu().numpy()), thr
        parts = [eng.preprofile_stage(batch, S, [-11.0, -1.0], threshold=thr, shard=(r, 2))[0].cpu().numpy() for r in range(2)]
        assert np.array_equal(parts[0] + parts[1], want.cpu().numpy()), thr
    asym = S.copy()
    asym[0, 1] += 1.0
    assert eng.dual_traced_ok(asym, [-11.0, -1.0], batch.lens) is None          # not symmetric: per-master path
    assert eng.dual_traced_ok(S, [-11.0, -1.0], batch.lens) is not None


@pytest.mark.parametrize("resident", ["one", "two"])
@pytest.mark.parametrize("mode", ["global", "local"])
def test_tensor_core_score_rows(eng, resident, mode):
    """k_build_rows_tc (tcgen05 tf32, hi/lo split) against the exact score rows and the CUDA-core
    tolerance kernel, element by element over a whole wave, then the alignment scores against the
    oracle within the stated 1e-5: several K classes (column chunks 32 .. 2 x 208), short row
    blocks, residents shorter than the tile, an asymmetric matrix."""
    rng = np.random.default_rng(8)
    S = matrices.blosum62().copy()
    S[3, 7] += 2.0                      # asymmetric: P.S and P.S^T differ
    lens = [31, 64, 65, 150, 300, 301, 410, 97, 5]
    profs = [synth.profile_from_counts(synth.count_profile(60 + k, L, 7 + k, 20, 27)) for k, L in enumerate(lens)]
    pb = eng.profile_batch(profs)
    pi, pj = synth.all_pairs(len(profs))
    rows = {}
    try:
        eng.keep_mwave = True
        for name, kw, tc in (("exact", dict(fast=False), True), ("fma", dict(fast=True), False), ("tc", dict(fast=True), True)):
            eng.fast_tc = tc
            sc = eng.align_profile_pairs(pb, pi, pj, S, [-11.0, -1.0], mode=mode, resident=resident, **kw)
            rows[name] = (eng.last_mwave.cpu().numpy().copy(), sc)
    finally:
        eng.keep_mwave, eng.fast_tc = False, True
    m_exact, s_exact = rows["exact"]
    m_tc, s_tc = rows["tc"]
    m_fma, _ = rows["fma"]
    fin = np.isfinite(m_exact)
    assert np.array_equal(np.isfinite(m_tc), fin) and np.array_equal(m_tc[~fin], m_exact[~fin])
    scale = np.abs(m_exact[fin]).max()
    assert np.abs(m_tc[fin] - m_exact[fin]).max() <= 4e-6 * scale
    assert np.abs(m_tc[fin] - m_fma[fin]).max() <= 4e-6 * scale
    assert np.allclose(s_tc, s_exact, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("mode", ["global", "local"])
def test_dense_profile_rows_packed_f32x2(eng, mode):
    """k_build_rows_x2 (packed f32x2, union of the resident's symbols, two streamed rows per register) gives the
    bits of k_build_rows_t over a whole wave of DENSE profiles -- and both the oracle's scores (cext.c:63-95 order):
    ragged lengths over several K classes (one above 512 columns: two column blocks), a sparse profile and a rare
    symbol in the batch (zero-padded union entries), odd row counts per block, an asymmetric matrix."""
    S = matrices.blosum62().copy()
    S[3, 7] += 2.0
    S[20:24, :20] = np.arange(-4, 0, dtype=np.float32)[:, None]
    lens = [5, 31, 64, 65, 150, 301, 410, 520, 97, 33]
    profs = [synth.profile_from_counts(synth.count_profile(160 + k, L, 150 + 300 * k, 20, 27)) for k, L in enumerate(lens)]
    profs[8] = synth.profile_from_counts(synth.count_profile(99, 97, 3, 20, 27))            # a sparse one
    c = synth.count_profile(98, 33, 400, 20, 27)
    c[7, 22] = 5                                                                              # a rare symbol
    profs[9] = synth.profile_from_counts(c)
    pb = eng.profile_batch(profs)
    assert pb.dense_syms() == 21
    pi, pj = synth.all_pairs(len(profs))
    pi, pj = np.concatenate([pi, pj[:12]]), np.concatenate([pj, pi[:12]])
    rows = {}
    try:
        eng.keep_mwave = True
        for x2 in (True, False):
            eng.rows_x2 = x2
            l0 = eng.launches
            sc = eng.align_profile_pairs(pb, pi, pj, S, [-11.0, -1.0], mode=mode, resident="one")
            rows[x2] = (eng.last_mwave.cpu().numpy().copy(), sc)
    finally:
        eng.keep_mwave, eng.rows_x2 = False, True
    assert np.array_equal(rows[True][0].view(np.uint32), rows[False][0].view(np.uint32))      # the last wave, bit for bit
    assert np.array_equal(rows[True][1].view(np.uint32), rows[False][1].view(np.uint32))
    for k in range(0, len(pi), 3):
        p1, p2 = profs[pi[k]], profs[pj[k]]
        g1, g2 = oracle.gap_arrays(p1.shape[0], p2.shape[0], [-11.0, -1.0])
        want, _ = oracle.align_raw(mode, oracle.build_scores([p1], [p2], [S]), g1, g2)
        assert float(rows[True][1][k]) == want, k
    # a small alphabet in use (dense DNA profiles: 4 of 15 symbols, two-entry table rows)
    Sn = matrices.nucleotide()
    dna = [synth.profile_from_counts(synth.count_profile(260 + k, L, 60, 4, 15)) for k, L in enumerate([40, 129, 64, 7])]
    pbn = eng.profile_batch(dna)
    assert pbn.dense_syms() == 4
    qi, qj = synth.all_pairs(len(dna))
    got = eng.align_profile_pairs(pbn, qi, qj, Sn, [-11.0, -1.0], mode=mode, resident="one")
    for k in range(len(qi)):
        p1, p2 = dna[qi[k]], dna[qj[k]]
        g1, g2 = oracle.gap_arrays(p1.shape[0], p2.shape[0], [-11.0, -1.0])
        want, _ = oracle.align_raw(mode, oracle.build_scores([p1], [p2], [Sn]), g1, g2)
        assert float(got[k]) == want, ("dna", k)
