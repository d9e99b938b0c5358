"""The numpy lane-model of the streaming kernel must agree with the oracle (CPU only)."""
import numpy as np
import pytest

import oracle
from praline_b200 import matrices, synth
from conftest import MODES
import model_stream as ms


def semiglobal_extend(path, L1, L2):
    path = [tuple(p) for p in path]
    pre = []
    if path[0][0] != 0:
        pre = [(i, 0) for i in range(path[0][0])]
    elif path[0][1] != 0:
        pre = [(0, i) for i in range(path[0][1])]
    post = []
    if path[-1][0] != L1:
        post = [(i, path[-1][1]) for i in range(path[-1][0] + 1, L1 + 1)]
    elif path[-1][1] != L2:
        post = [(path[-1][0], i) for i in range(path[-1][1] + 1, L2 + 1)]
    return pre + path + post


def finish(mode, transposed, o, info, tb, K, L_res, L_str):
    """score + reference-orientation path from the model's per-pair outputs."""
    B = info["B"]

    def code_at(yk, xk):
        if yk == 0 and xk == 0:
            return B["code00"]
        if yk == 0:
            return 2
        if xk == 0:
            return 1
        nb = ms.fetch_nib(tb, K, o["emit_t"], info["lr"], L_str, yk, xk)
        if not nb & 1:
            return 0
        if transposed:
            return 1 if nb & 2 else 2
        return 2 if nb & 2 else 1

    if mode == 0:
        start = (L_str, L_res, code_at(L_str, L_res))
        score = o["score"]
    else:
        rowk, colk = o["rowkey"], o["colkey"]
        from_row = mode in (2, 4)
        if not transposed:
            if rowk[0] > colk[0] and from_row:
                score, cell = rowk[0], (L_str, rowk[1])
            else:
                score, cell = colk[0], (colk[1], L_res)
        else:
            if colk[0] > rowk[0] and from_row:
                score, cell = colk[0], (colk[1], L_res)
            else:
                score, cell = rowk[0], (L_str, rowk[1])
        start = (cell[0], cell[1], code_at(*cell))
    path = ms.traceback(tb, K, info, L_str, L_res, o["emit_t"], start, transposed)
    if mode != 0:
        L1, L2 = (L_res, L_str) if transposed else (L_str, L_res)
        path = semiglobal_extend(path, L1, L2)
    return float(score), path


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("mode", [0, 2, 3, 4])
def test_model_paths_match_oracle(mode, transposed):
    S = matrices.blosum62()
    rng = np.random.default_rng(100 + mode + 10 * transposed)
    for trial, (K, gaps) in enumerate([(2, (-11.0, -1.0)), (3, (-8.0, -8.0)), (5, (-1.0, -1.0)), (2, (-4.0, -2.0))]):
        fam = synth.family(50 + trial + mode, 6, int(rng.integers(20, 32 * K - 8)))
        resident = fam[0]
        streams = fam[1:] + [rng.integers(0, 20, 7).astype(np.int32), fam[2][:1]]
        out, tb, info = ms.run_warp(K, resident, streams, S, gaps[0], gaps[1], mode, transposed, want_tb=True)
        for s, o in zip(streams, out):
            one, two = (resident, s) if transposed else (s, resident)
            want_score, want_path = oracle.align_seqs(MODES[mode], one, two, S, list(gaps))
            score, path = finish(mode, transposed, o, info, tb, K, len(resident), len(s))
            assert score == want_score, (mode, transposed, trial)
            assert path == [tuple(p) for p in want_path.tolist()], (mode, transposed, trial)


@pytest.mark.parametrize("transposed", [False, True])
def test_model_local_scores_match_oracle(transposed):
    S = matrices.blosum62()
    fam = synth.family(77, 5, 40)
    out, tb, info = ms.run_warp(2, fam[0], fam[1:], S, -11.0, -1.0, 1, transposed)
    for s, o in zip(fam[1:], out):
        one, two = (fam[0], s) if transposed else (s, fam[0])
        want, _ = oracle.align_seqs("local", one, two, S, [-11.0, -1.0])
        assert float(o["score"]) == want


def test_local_model_matches_oracle_with_boxes():
    """The local traced encoding (tests/model_local.py = what gotoh_stream.cuh / traceback.cu do in
    local mode) gives the oracle's scores, paths, boxes and preprofile counts over Waterman-Eggert
    iterations, ties on W-W = -gap_open and linear gaps included."""
    import model_local
    import we_model
    from praline_b200 import matrices, synth
    S = matrices.blosum62()
    rng = np.random.default_rng(5)
    fam = [np.asarray(s) for s in synth.family(17, 4, 36)] + [rng.integers(0, 20, 15).astype(np.int32)]
    w = 17   # tryptophan: S[W][W] = 11 ties a gap opening of -11 against the local zero
    fam.append(np.array([w, 3, w, w, 5, w, 2, w, w], np.int32))
    fam.append(np.array([w, w, 4, w, 6, 6, w, w, 1, w], np.int32))
    for gaps in ([-11.0, -1.0], [-3.0], [-11.0, 0.0]):
        go, ge = (gaps[0], gaps[-1])
        for i, a in enumerate(fam):
            want_counts, _, _ = we_model.local_master_counts(fam, i, S, gaps, 3, None, 27)
            counts = np.zeros_like(want_counts)
            counts[np.arange(len(a)), a] += 1
            for j, b in enumerate(fam):
                if i == j:
                    continue
                boxes = []
                for K in (1, 3):
                    boxes = []
                    for score, path, _ in we_model.we_alignments(a, b, S, gaps, 3):
                        nib, key = model_local.fill(a, b, S, go, ge, boxes, K=K)
                        got = model_local.walk(nib, key, boxes)
                        assert key[0] == score, (gaps, i, j)
                        assert np.array_equal(got, path), (gaps, i, j, len(boxes))
                        boxes.append(we_model.boxes_of([path])[0])
                        if K == 3:
                            model_local.counts_from_path(got, b, len(a), 27, counts)
            assert np.array_equal(counts, want_counts), (gaps, i)
