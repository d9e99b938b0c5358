"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the host
tiling logic (no kernel is launched here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from praline_b200 import _lib, synth
from praline_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "praline_b200.h")).read()
    declared = set(re.findall(r"\b(pgpu_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().pgpu_abi_version() == 1


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.PralineGpuError):
        E.Engine(0)
    assert _lib.load().pgpu_init(0) != 0
    assert b"CUDA" in _lib.load().pgpu_last_error() or b"device" in _lib.load().pgpu_last_error()


def test_gap_series_rules():
    assert E._gaps([-8.0]) == (np.float32(-8), np.float32(-8))
    assert E._gaps([-11.0, -1.0]) == (np.float32(-11), np.float32(-1))
    with pytest.raises(ValueError):
        E._gaps([-11.0, -2.0, -1.0])


def test_borders_match_reference_init():
    import oracle
    for mode, name in enumerate(["global", "local", "semiglobal_both", "semiglobal_one", "semiglobal_two"]):
        g1, g2 = oracle.gap_arrays(40, 40, [-11.0, -1.0])
        o, t = oracle.ref_init_borders(name, g1, g2, 40, 40)
        B = E.borders(mode, -11.0, -1.0, 40, transposed=False)
        assert np.array_equal(B["topD"], o[0].max(axis=1))
        assert np.array_equal(B["leftD"], o[:, 0].max(axis=1))
        Bt = E.borders(mode, -11.0, -1.0, 40, transposed=True)
        assert np.array_equal(Bt["topD"], B["leftD"]) and np.array_equal(Bt["leftD"], B["topD"])
        assert B["top_ramp"] == int(t[0, 1, 2] != 0) and B["left_ramp"] == int(t[1, 0, 1] != 0)


class _FakeEngine(E.Engine):
    def __init__(self):
        self.nw = 8
        self.k_set = [1, 2, 3, 4, 6, 8, 10, 12, 13, 14, 16, 20, 24, 32]
        self.pin = False


def test_tiles_cover_every_pair_once():
    eng = _FakeEngine()
    rng = np.random.default_rng(0)
    res = np.sort(rng.integers(0, 30, 5000))
    tiles = eng._make_tiles(res, len(res), 48)
    seen = np.zeros(len(res), int)
    for t in tiles:
        assert t["stream_end"] - t["stream_begin"] <= 48 and t["out_base"] == t["stream_begin"]
        assert (res[t["stream_begin"]:t["stream_end"]] == t["resident"]).all()
        seen[t["stream_begin"]:t["stream_end"]] += 1
    assert (seen == 1).all()


def test_allpairs_tiles_partition_condensed_vector():
    eng = _FakeEngine()

    class B(object):
        pass
    for n, world in [(2, 1), (37, 1), (37, 4), (200, 8)]:
        b = B()
        b.n = n
        b.lens = np.random.default_rng(n).integers(20, 400, n).astype(np.int64)
        pi, pj = synth.all_pairs(n)
        seen = np.zeros(n * (n - 1) // 2, int)
        tot = 0
        prev_hi = 0
        for r in range(world):
            by_k, (lo, hi), cells, cuts, _, _ = eng.allpairs_tiles(b, (r, world), tile=16)
            assert lo == prev_hi
            prev_hi = hi
            tot += cells
            for K, tiles in by_k.items():
                for t in tiles:
                    i = t["resident"]
                    assert 32 * K >= b.lens[i]
                    for s in range(t["stream_begin"], t["stream_end"]):
                        slot = t["out_base"] + s - t["stream_begin"]
                        assert pi[slot] == i and pj[slot] == s and lo <= slot < hi
                        seen[slot] += 1
        assert prev_hi == len(seen) and (seen == 1).all()
        assert tot == int((b.lens[pi] * b.lens[pj]).sum())


def test_tb_words_formula():
    eng = _FakeEngine()
    lens = np.asarray([5, 7, 3, 9, 4, 4, 6, 8, 2, 10, 11], np.int64)
    cs = np.concatenate([[0], np.cumsum(lens)])
    tiles = np.zeros(1, E.TILE_DTYPE)
    tiles["stream_begin"], tiles["stream_end"] = 0, 11
    w = eng._tb_words(tiles, 3, cs)[0]
    per = 2
    for k in range(8):
        sb, se = min(k * per, 11), min(k * per + per, 11)
        rows = lens[sb:se].sum()
        T = (rows + 1 + 31 + 31) // 32 * 32 if se > sb else 0
        assert w[k] == T // 8 * 3 * 32


def test_paired_allpairs_tiles_cover_every_pair_once():
    """Paired-resident plan of the packed int16 kernel: slots of resident A and resident B."""
    eng = _FakeEngine()

    class B(object):
        pass
    for n, world in [(2, 1), (3, 1), (38, 1), (37, 3), (120, 8)]:
        b = B()
        b.n = n
        b.lens = np.random.default_rng(n).integers(20, 400, n).astype(np.int64)
        pi, pj = synth.all_pairs(n)
        seen = np.zeros(n * (n - 1) // 2, int)
        prev_hi = 0
        for r in range(world):
            by_k, (lo, hi), cells, cuts, _, paired = eng.allpairs_tiles(b, (r, world), tile=16, paired=True)
            assert paired and lo == prev_hi
            prev_hi = hi
            for K, tiles in by_k.items():
                for t in tiles:
                    i, i2 = int(t["resident"]), int(t["resident2"])
                    assert 32 * K >= b.lens[i] and (i2 < 0 or (i2 == i + 1 and 32 * K >= b.lens[i2]))
                    for s in range(t["stream_begin"], t["stream_end"]):
                        e = s - t["stream_begin"]
                        slot = t["out_base"] + e
                        assert pi[slot] == i and pj[slot] == s and lo <= slot < hi
                        seen[slot] += 1
                        if i2 >= 0 and e >= t["b_skip"]:
                            slot2 = t["out_base2"] + e - t["b_skip"]
                            assert pi[slot2] == i2 and pj[slot2] == s and lo <= slot2 < hi
                            seen[slot2] += 1
                        elif i2 >= 0:
                            assert s == i2          # the skipped element is resident2 itself
        assert prev_hi == len(seen) and (seen == 1).all()


def test_profile_wave_blocks_cover_rows():
    """Row blocks of a profile wave: every matrix row exactly once, fed by the right profile row."""
    rng = np.random.default_rng(0)
    nseq = 9
    lens = rng.integers(3, 80, nseq).astype(np.int64)
    offs = np.r_[0, np.cumsum(lens)]
    n = 40
    res_s = np.sort(rng.integers(0, nseq, n))
    str_s = rng.integers(0, nseq, n)
    eng = _FakeEngine()
    tiles = eng._make_tiles(res_s, n, 16)
    lens_s = lens[str_s]
    cs = np.r_[0, np.cumsum(lens_s)]
    mb, blocks, nr = E.plan_profile_wave(tiles, 8, cs, lens_s, str_s, res_s, offs)
    src = np.full(nr, -9, np.int64)
    res = np.full(nr, -9, np.int64)
    for b in blocks:
        assert 1 <= b["rows"] <= 32
        for r in range(b["rows"]):
            assert src[b["row0"] + r] == -9
            src[b["row0"] + r] = -1 if (b["dummy"] and r == 0) else b["src0"] + r
            res[b["row0"] + r] = b["res"]
    for ti, t in enumerate(tiles):
        bb, e = int(t["stream_begin"]), int(t["stream_end"])
        per = (e - bb + 7) // 8
        for w in range(8):
            sb = min(bb + w * per, e)
            se = min(sb + per, e)
            if se <= sb:
                continue
            r = mb[ti * 8 + w]
            assert src[r] == -1 and res[r] == t["resident"]
            r += 1
            for el in range(sb, se):
                for y in range(lens_s[el]):
                    assert src[r] == offs[str_s[el]] + y and res[r] == res_s[el]
                    r += 1
    assert (src != -9).all()


def test_tb_words16_formula():
    """Traceback words of the packed traced kernel: a warp's slice runs as two halves in lock step
    (gotoh_stream16.cu: mid = sb + (se - sb + 1) / 2), four steps per word."""
    eng = _FakeEngine()
    rng = np.random.default_rng(4)
    lens = rng.integers(1, 60, 37).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(lens)])
    tiles = np.zeros(2, E.TILE_DTYPE)
    tiles["stream_begin"], tiles["stream_end"] = [0, 20], [20, 37]
    w = eng._tb_words16(tiles, 5, cs)
    for ti, (b, e) in enumerate([(0, 20), (20, 37)]):
        per = (e - b + 7) // 8
        for k in range(8):
            sb, se = min(b + k * per, e), min(b + k * per + per, e)
            mid = sb + (se - sb + 1) // 2
            rows = max(lens[sb:mid].sum(), lens[mid:se].sum())
            T = (rows + 1 + 31 + 31) // 32 * 32 if se > sb else 0
            assert w[ti, k] == T // 4 * 5 * 32


def test_k_classes_match_k_for():
    eng = _FakeEngine()
    lens = np.asarray([1, 31, 32, 33, 64, 65, 150, 319, 320, 321, 400, 416, 417, 1000, 1024, 1025, 5000])
    got = eng.k_classes(lens)
    want = [eng.k_for(int(l)) or -1 for l in lens]
    assert got.tolist() == want


def test_fits_s16_ranges():
    """The traced packed kernel needs every DP value inside +-16000 (its tie tests subtract two
    values), the score-only one inside +-32000; non-integer scores never qualify."""
    from praline_b200 import matrices
    eng = _FakeEngine()
    S = matrices.blosum62()
    assert eng.fits_s16(S, -11.0, -1.0, np.asarray([300, 300])) is not None
    assert eng.fits_s16(S, -11.0, -1.0, np.asarray([300, 400]), limit=16000) is not None
    assert eng.fits_s16(S, -11.0, -1.0, np.asarray([900]), limit=16000) is None
    assert eng.fits_s16(S, -11.0, -1.0, np.asarray([900])) is not None
    assert eng.fits_s16(S, -11.0, -1.0, np.asarray([2000])) is None
    assert eng.fits_s16(S, -11.5, -1.0, np.asarray([100])) is None
    assert eng.fits_s16(S * 0.5, -11.0, -1.0, np.asarray([100])) is None


def test_row_block_quads():
    """128-row tiles of the tensor-core score rows: <= 4 consecutive row blocks of one resident."""
    from praline_b200.engine import row_block_quads, ROWBLOCK_DTYPE
    b = np.zeros(11, ROWBLOCK_DTYPE)
    b["res"] = [5, 5, 5, 5, 5, 5, 7, 7, 9, 9, 9]
    assert row_block_quads(b).tolist() == [[0, 4], [4, 2], [6, 2], [8, 3]]
    assert row_block_quads(b[:0]).shape == (0, 2)
    rng = np.random.default_rng(0)
    res = np.repeat(rng.integers(0, 50, 300), rng.integers(1, 12, 300))
    b = np.zeros(len(res), ROWBLOCK_DTYPE)
    b["res"] = res
    q = row_block_quads(b)
    assert q[:, 1].sum() == len(res) and (q[:, 1] >= 1).all() and (q[:, 1] <= 4).all()
    assert (q[1:, 0] == q[:-1, 0] + q[:-1, 1]).all()
    for f, c in q:
        assert len(set(res[f:f + c])) == 1


def test_row_block_quad_descriptors():
    """pgpu_quad records: one 128-byte line per tile with the resident's rows and every block's fields."""
    from praline_b200.engine import row_block_quads, ROWBLOCK_DTYPE, QUAD_DTYPE
    assert QUAD_DTYPE.itemsize == 160
    b = np.zeros(7, ROWBLOCK_DTYPE)
    b["res"] = [2, 2, 2, 2, 2, 0, 0]
    b["row0"] = [0, 32, 64, 96, 128, 160, 170]
    b["src0"] = [99, 131, 163, 195, 227, 5, 15]
    b["rows"] = [32, 32, 32, 32, 7, 10, 3]
    b["dummy"] = [1, 0, 0, 0, 0, 1, 0]
    offs = np.array([0, 40, 100, 400], np.int64)
    q = row_block_quads(b, offs)
    assert q["nblk"].tolist() == [4, 1, 2] and q["q0"].tolist() == [100, 100, 0] and q["Lr"].tolist() == [300, 300, 40]
    assert q["row0"].tolist() == [[0, 32, 64, 96], [128, 0, 0, 0], [160, 170, 0, 0]]
    assert q["src0"].tolist() == [[99, 131, 163, 195], [227, 0, 0, 0], [5, 15, 0, 0]]
    assert q["rows"].tolist() == [[32, 32, 32, 32], [7, 0, 0, 0], [10, 3, 0, 0]]
    assert q["dummy"].tolist() == [[1, 0, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]]
    assert (q["can0"] == -1).all()                       # no pre-split store given
    # with a pre-split store: blocks that start on an 8-row group of their streamed sequence get a TMA row
    padoff = np.array([0, 64, 128, 448], np.int64)
    b["src0"] = [39, 40, 72, 99, 100, 132, 5]           # seq 0 rows 39 (dummy in front of seq 1), seq 1 rows 0, 32, 59; seq 2 rows 0, 32; seq 0 row 5
    q = row_block_quads(b, offs, padoff)
    assert q["bcan"].tolist() == [128, 128, 0]
    assert q["can0"].tolist() == [[-1, 64, 96, -1], [128, -1, -1, -1], [-1, -1, -1, -1]]
    assert len(row_block_quads(b[:0], offs)) == 0


def test_tf32_split_accuracy_model():
    """Numerical model of K1t's FP32-accurate split (score_rows_tc.cu): hi = cvt.rna.tf32(x), lo = x - hi (exact in
    f32), the tensor core reads the top 19 bits of every operand, products are exact and summed in f32.  The three
    kept terms hi.hi + lo.hi + hi.lo stay within 2e-6 of sum |terms| of the exact contraction (stated bound 1e-5)."""
    rng = np.random.default_rng(4)

    def tf32_rna(x):        # round to nearest, ties away, 10 explicit mantissa bits
        b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
        b = (b + 0x1000) & 0xffffe000
        return b.astype(np.uint32).view(np.float32)

    def tf32_trunc(x):      # what the MMA does with a 32-bit operand register
        return (x.astype(np.float32).view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)

    A = 27
    P = rng.dirichlet(np.ones(A) * 0.3, size=400).astype(np.float32)           # profile rows
    S = rng.integers(-4, 12, (A, A)).astype(np.float32)
    W = (rng.dirichlet(np.ones(A) * 0.3, size=300).astype(np.float32) @ S.T).astype(np.float32)
    Ph, Wh = tf32_rna(P), tf32_rna(W)
    Pl, Wl = tf32_trunc(P - Ph), tf32_trunc(W - Wh)
    assert np.array_equal(Ph + (P - Ph), P) and np.array_equal(Wh + (W - Wh), W)      # the split itself is exact
    got = (Ph.astype(np.float64) @ Wh.astype(np.float64).T + Pl.astype(np.float64) @ Wh.astype(np.float64).T
           + Ph.astype(np.float64) @ Wl.astype(np.float64).T).astype(np.float32)
    exact = P.astype(np.float64) @ W.astype(np.float64).T
    scale = np.abs(P).astype(np.float64) @ np.abs(W).astype(np.float64).T
    assert (np.abs(got - exact) <= 2e-6 * scale).all()
    one_term = (Ph.astype(np.float64) @ Wh.astype(np.float64).T)
    assert (np.abs(one_term - exact) > 1e-5 * scale).any()       # plain tf32 would NOT meet the bound


def test_dual_wave_plan_covers_every_pair_once():
    """Host plan of the symmetric traced all-vs-all (engine.DualWavePlan, pure numpy): over all waves and both
    ranks of a 2-way shard every unordered pair appears exactly once, slots of the two residents interleave without
    collisions, word regions of (tile, warp) do not overlap, no wave exceeds the traceback budget (beyond one tile),
    a wave holds one columns-per-lane class, and the cell count matches."""
    from praline_b200 import engine as E
    rng = np.random.default_rng(11)
    for n, tile, budget, chunk in ((2, 16, 1 << 12, 3), (3, 16, 1 << 12, 1), (37, 16, 1 << 14, 2), (64, 8, 1 << 13, 5)):
        lens = rng.integers(1, 140, n).astype(np.int64)
        lens[rng.integers(0, n)] = 300           # a second K class
        ks = np.array([1, 2, 3, 4, 6, 8, 10, 12, 13, 14, 16, 20, 24, 32])
        kcls = ks[np.searchsorted(32 * ks, lens, side="left")]
        seen = set()
        cells = 0
        for rank in range(2):
            plan = E.DualWavePlan(lens, kcls, 8, tile, budget, (rank, 2), chunk=chunk)
            cells += plan.cells
            for K, wt, wbase, ns, n_words, na, nb in plan.waves():
                words = E.tb_words16r(wt, K, plan.cs, 8)
                assert n_words == int(words.sum()) and np.array_equal(wbase, np.concatenate([[0], np.cumsum(words.ravel())[:-1]]))
                assert n_words <= budget or len(wt) == 1
                slots = set()
                for t in wt:
                    r1, r2, b0, b1, sk, st = int(t["resident"]), int(t["resident2"]), int(t["stream_begin"]), int(t["stream_end"]), int(t["b_skip"]), int(t["reserved"])
                    assert st == 2 and K == max(kcls[r1], kcls[r2] if r2 >= 0 else 0)
                    for e in range(b1 - b0):
                        j = b0 + e
                        sa = int(t["out_base"]) + e * st
                        assert r1 < j and (r1, j) not in seen and sa not in slots and sa < ns
                        seen.add((r1, j)); slots.add(sa)
                        if r2 >= 0 and e >= sk:
                            sb_ = int(t["out_base2"]) + (e - sk) * st
                            assert r2 < j and (r2, j) not in seen and sb_ not in slots and sb_ < ns
                            seen.add((r2, j)); slots.add(sb_)
        assert seen == {(i, j) for i in range(n) for j in range(i + 1, n)}
        ii, jj = np.triu_indices(n, 1)
        assert cells == int((lens[ii] * lens[jj]).sum())


def test_dense_symbol_count():
    """Host side of the packed f32x2 score rows (Engine.align_profile_pairs -> pgpu_build_rows): the number of symbols
    in use is passed only for dense batches; sparse ones and large alphabets stay on k_build_rows_t."""
    from praline_b200.engine import dense_symbol_count
    used = np.zeros(27, bool)
    used[:20] = True
    used[22] = True
    assert dense_symbol_count(used, int(0.9 * 21 * 1000), 1000, 27) == 21
    assert dense_symbol_count(used, int(0.3 * 21 * 1000), 1000, 27) == 0
    assert dense_symbol_count(np.ones(40, bool), 10 ** 6, 100, 40) == 0
    assert dense_symbol_count(np.zeros(27, bool), 0, 100, 27) == 0


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_native_profile_wave_planner_equals_numpy(seed):
    """pgpu_plan_profile_wave (csrc/host_plan.cu, host code of the library) against its numpy specification
    plan_profile_wave + row_block_quads: region rows, row blocks and quads field by field, with and without the
    pre-split row offsets, ragged lengths incl. multiples of 32 and single rows, empty warp regions."""
    rng = np.random.default_rng(seed)
    nseq = 11
    lens = rng.choice([1, 2, 31, 32, 33, 64, 65, 100, 7], nseq).astype(np.int64)
    offs = np.r_[0, np.cumsum(lens)].astype(np.int64)
    padoff = np.r_[0, np.cumsum((lens + 31) // 32 * 32)].astype(np.int64)
    n = int(rng.integers(3, 70))
    res_s = np.sort(rng.integers(0, nseq, n)).astype(np.int64)
    str_s = rng.integers(0, nseq, n).astype(np.int64)
    eng = _FakeEngine()
    tiles = eng._make_tiles(res_s, n, int(rng.integers(2, 20)))
    lens_s = lens[str_s]
    cs = np.r_[0, np.cumsum(lens_s)].astype(np.int64)
    for nw in (8, 3):
        mb, blocks, nr = E.plan_profile_wave(tiles, nw, cs, lens_s, str_s, res_s, offs)
        mb2, blocks2, nr2 = E.plan_profile_wave_native(tiles, nw, cs, lens_s, str_s, res_s, offs)
        assert nr == nr2 and np.array_equal(mb, mb2)
        assert blocks2.dtype == blocks.dtype and len(blocks) == len(blocks2)
        for f in ("row0", "src0", "rows", "res", "dummy"):
            assert np.array_equal(blocks[f], blocks2[f]), f
        for pad in (None, padoff):
            want = E.row_block_quads(blocks, offs, pad)
            mb3, blocks3, nr3, quads = E.plan_profile_wave_native(tiles, nw, cs, lens_s, str_s, res_s, offs,
                                                                  want_quads=True, padoff=pad)
            assert nr3 == nr and np.array_equal(mb3, mb) and np.array_equal(blocks3["row0"], blocks["row0"])
            assert len(quads) == len(want)
            for f in ("q0", "Lr", "nblk", "row0", "src0", "rows", "dummy", "bcan", "can0"):
                assert np.array_equal(quads[f], want[f]), (f, pad is None)
