#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs the reference installed in baseline/_ref,
see DESIGN.md: `pip install --no-deps --target baseline/_ref` from a copy of
/root/reference).  Every vector is produced through the reference's public
component API: Manager -> PairwiseAligner -> RawPairwiseAligner
(praline/component/align.py:37-447), i.e. cext_build_scores + cext_align_* +
end-cell choice + get_paths + extend_path_semiglobal.

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/matrices.npz        blosum62 [27x27], nucleotide [15x15] as the reference loads them
    tests/golden/pairwise_seq.json   sequence-sequence cases: inputs, score, path, path container type
    tests/golden/pairwise_prof.npz   profile-profile cases: counts, m, score, path
    tests/golden/fill_cells.npz      full o/t matrices of small cases from cext_align_* (all 5 modes)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests", "_stubs"))
sys.path.insert(0, ROOT)

import warnings
warnings.filterwarnings("ignore")

from praline import load_score_matrix, open_builtin  # noqa: E402
from praline.core import TypeIndex, Manager, Execution, Environment  # noqa: E402
ROOT_TAG = "__ROOT_TAG__"  # praline/cmd.py:22
from praline.container import (Sequence, PlainTrack, ProfileTrack, ALPHABET_AA, ALPHABET_DNA,  # noqa: E402
                               TRACK_ID_INPUT, TRACK_ID_PREPROFILE, MatchScoreModel, GapScoreModel)
from praline.component import PairwiseAligner, RawPairwiseAligner  # noqa: E402
from praline.util import cext_build_scores  # noqa: E402
import praline.util as putil  # noqa: E402

from praline_b200 import synth  # noqa: E402

MODES = ["global", "local", "semiglobal_both", "semiglobal_one", "semiglobal_two"]


def manager():
    index = TypeIndex()
    index.register(PairwiseAligner)
    index.register(RawPairwiseAligner)
    return Manager(index)


def run_pair(mgr, mode, seq1, seq2, trid, sm, gaps, zero_idxs=None):
    env = Environment(keys={'gap_series': [float(g) for g in gaps]})
    ex = Execution(mgr, ROOT_TAG)
    task = ex.add_task(PairwiseAligner)
    task.environment(env)
    kw = dict(mode=mode, sequence_one=seq1, sequence_two=seq2, track_id_sets_one=[[trid]],
              track_id_sets_two=[[trid]], score_matrices=[sm])
    if zero_idxs is not None:
        kw['zero_idxs'] = zero_idxs
    task.inputs(**kw)
    for _ in ex.run():
        pass
    out = ex.outputs[0]
    path = out['alignment'].path
    ptype = "ndarray" if isinstance(path, np.ndarray) else "list"
    return float(out['score']), np.asarray(path).astype(int).tolist(), ptype


def main():
    mgr = manager()
    with open_builtin('matrices/blosum62') as f:
        blosum = load_score_matrix(f, alphabet=ALPHABET_AA)
    with open_builtin('matrices/nucleotide') as f:
        nuc = load_score_matrix(f, alphabet=ALPHABET_DNA)
    np.savez(os.path.join(HERE, "matrices.npz"), blosum62=blosum.matrix.astype(np.float32),
             nucleotide=nuc.matrix.astype(np.float32))

    # ---- sequence-sequence --------------------------------------------------------------
    cases = []

    def aa(s):
        return Sequence("s", [(TRACK_ID_INPUT, PlainTrack(s, ALPHABET_AA))])

    def add(name, a_idx, b_idx, mode, gaps, zero=None, alphabet="aa"):
        alpha = ALPHABET_AA if alphabet == "aa" else ALPHABET_DNA
        sm = blosum if alphabet == "aa" else nuc
        s1 = Sequence("a", [(TRACK_ID_INPUT, PlainTrack(None, alpha, raw_indices=np.asarray(a_idx)))])
        s2 = Sequence("b", [(TRACK_ID_INPUT, PlainTrack(None, alpha, raw_indices=np.asarray(b_idx)))])
        score, path, ptype = run_pair(mgr, mode, s1, s2, TRACK_ID_INPUT, sm, gaps, zero)
        cases.append(dict(name=name, alphabet=alphabet, a=[int(v) for v in a_idx], b=[int(v) for v in b_idx],
                          mode=mode, gaps=[float(g) for g in gaps],
                          zero_idxs=None if zero is None else [list(map(int, z)) for z in zero],
                          score=score, path=path, path_type=ptype))

    # SURVEY.md section 4 known-answer table
    a = [ALPHABET_AA.symbol_to_index(c) for c in "HEAGAWGHEE"]
    b = [ALPHABET_AA.symbol_to_index(c) for c in "PAWHEAE"]
    for mode in MODES:
        for gaps in ([-11.0, -1.0], [-8.0]):
            add("kat", a, b, mode, gaps)
    box = [(y, x) for y in range(1, 6) for x in range(2, 8)]
    add("kat_masked", a, b, "local", [-11.0, -1.0], zero=box)

    # seeded families: short/ragged, all modes, affine + linear + odd gap values
    rng = np.random.default_rng(1234)
    fam = synth.family(11, 12, 40)
    fam += synth.family(12, 6, 9)
    fam += [np.asarray([3], np.int32), np.asarray([3, 3], np.int32), rng.integers(0, 20, 70).astype(np.int32)]
    fam += synth.family(13, 4, 130)
    for k in range(60):
        i, j = rng.integers(0, len(fam), 2)
        mode = MODES[k % 5]
        gaps = [[-11.0, -1.0], [-8.0], [-4.0, -2.0], [-1.0, -1.0], [-2.5, -0.5], [-6.0, -6.0]][k % 6]
        add("fam%d" % k, fam[i], fam[j], mode, gaps)
    # sequences that use the ambiguity symbols B Z X * and the unscored U O J (zero rows)
    amb = rng.integers(0, 27, 50).astype(np.int32)
    amb2 = rng.integers(0, 27, 44).astype(np.int32)
    for mode in MODES:
        add("ambig", amb, amb2, mode, [-11.0, -1.0])
    # DNA
    dfam = synth.family(14, 4, 60, n_sym=4)
    for k, mode in enumerate(MODES):
        add("dna%d" % k, dfam[k % 4], dfam[(k + 1) % 4], mode, [-11.0, -1.0] if k % 2 else [-2.0], alphabet="dna")
    # Waterman-Eggert style masks on a longer local alignment
    big_a, big_b = fam[0], fam[1]
    zero = [(y, x) for y in range(5, 20) for x in range(4, 22)]
    add("masked_fam", big_a, big_b, "local", [-11.0, -1.0], zero=zero)
    add("masked_fam_lin", big_a, big_b, "local", [-8.0], zero=zero)

    with open(os.path.join(HERE, "pairwise_seq.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("sequence cases:", len(cases))

    # ---- profile-profile -----------------------------------------------------------------
    prof = {}
    pc = 0
    for k in range(12):
        alpha, sm, nsym, A = (ALPHABET_AA, blosum, 20, 27) if k % 3 else (ALPHABET_DNA, nuc, 4, 15)
        L1, L2 = [(23, 31), (40, 40), (17, 55), (64, 50)][k % 4]
        c1 = synth.count_profile(100 + k, L1, 3 + k, nsym, A)
        c2 = synth.count_profile(200 + k, L2, 2 + (k * 5) % 7, nsym, A)
        mode = MODES[k % 5]
        gaps = [[-11.0, -1.0], [-2.0], [-5.5, -0.75]][k % 3]
        s1 = Sequence("p1", [(TRACK_ID_PREPROFILE, ProfileTrack(c1, alpha))])
        s2 = Sequence("p2", [(TRACK_ID_PREPROFILE, ProfileTrack(c2, alpha))])
        score, path, ptype = run_pair(mgr, mode, s1, s2, TRACK_ID_PREPROFILE, sm, gaps)
        p1 = s1.get_track(TRACK_ID_PREPROFILE).profile.astype(np.float32)
        p2 = s2.get_track(TRACK_ID_PREPROFILE).profile.astype(np.float32)
        from praline.component.align import build_nonzero_matrix
        m = np.zeros((L1, L2), np.float32)
        cext_build_scores([p1], [p2], [build_nonzero_matrix(p1)], [build_nonzero_matrix(p2)],
                          [sm.matrix.astype(np.float32)], m)
        pre = "c%d_" % pc
        prof[pre + "counts1"] = c1
        prof[pre + "counts2"] = c2
        prof[pre + "alphabet"] = np.asarray(A)
        prof[pre + "mode"] = np.asarray(MODES.index(mode))
        prof[pre + "gaps"] = np.asarray(gaps, np.float64)
        prof[pre + "m"] = m
        prof[pre + "score"] = np.asarray(score, np.float64)
        prof[pre + "path"] = np.asarray(path, np.int32)
        pc += 1
    prof["n"] = np.asarray(pc)
    np.savez_compressed(os.path.join(HERE, "pairwise_prof.npz"), **prof)
    print("profile cases:", pc)

    # ---- full o / t matrices straight from cext_align_* ------------------------------------
    cells = {}
    cc = 0
    for k, mode in enumerate(MODES):
        for gaps in ([-11.0, -1.0], [-3.0]):
            a_idx, b_idx = fam[k], fam[k + 5]
            S = blosum.matrix.astype(np.float32)
            m = S[np.asarray(a_idx)][:, np.asarray(b_idx)].astype(np.float32).copy()
            if k % 2:  # non-integer match scores too
                m = (m * np.float32(0.37)).astype(np.float32)
            gs = gaps if len(gaps) == 2 else [gaps[0], gaps[0]]
            g1 = np.empty((m.shape[0], 2), np.float32); g1[:] = gs
            g2 = np.empty((m.shape[1], 2), np.float32); g2[:] = gs
            zero = [(3, 4), (3, 5), (4, 4), (10, 10)] if cc % 3 == 0 else None
            # drive RawPairwiseAligner's own array prep by re-stating it through the component:
            s1 = Sequence("a", [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=np.asarray(a_idx)))])
            s2 = Sequence("b", [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=np.asarray(b_idx)))])
            shape = (m.shape[0] + 1, m.shape[1] + 1)
            o = np.zeros(shape + (3,), np.float32)
            t = np.zeros(shape + (3,), np.uint8)
            z = np.zeros(shape, np.uint8)
            if zero:
                for idx in zero:
                    z[idx] = 1
            # praline/component/align.py:367-385, executed verbatim-equivalent via numpy
            o[:, 0, :] = -np.inf
            o[0, :, :] = -np.inf
            o[0, 0, 0] = 0
            if mode in {"semiglobal_both", "semiglobal_one"}:
                o[:, 0, 1] = 0
            else:
                o[0, 0, 1] = g1[0, 0] - g1[0, 1]
                o[1:, 0, 1] = (np.arange(o.shape[0] - 1) * g1[:, 1]) + g1[0, 0]
                t[1:, 0, 1] = putil.TRACEBACK_INSERT_UP_EXTEND
            if mode in {"semiglobal_both", "semiglobal_two"}:
                o[0, :, 2] = 0
            else:
                o[0, 0, 2] = g2[0, 0] - g2[0, 1]
                o[0, 1:, 2] = (np.arange(o.shape[1] - 1) * g2[:, 1]) + g2[0, 0]
                t[0, 1:, 2] = putil.TRACEBACK_INSERT_LEFT_EXTEND
            getattr(putil, "cext_align_" + mode)(m, g1, g2, o, t, z)
            # and the Raw component's answer on the same models
            ex = Execution(mgr, ROOT_TAG)
            task = ex.add_task(RawPairwiseAligner)
            task.environment(Environment(keys={}))
            task.inputs(mode=mode, sequence_one=s1, sequence_two=s2,
                        match_score_model=MatchScoreModel(s1, s2, m),
                        gap_score_model_one=GapScoreModel(s1, g1), gap_score_model_two=GapScoreModel(s2, g2),
                        zero_idxs=zero)
            for _ in ex.run():
                pass
            out = ex.outputs[0]
            pre = "c%d_" % cc
            cells[pre + "mode"] = np.asarray(MODES.index(mode))
            cells[pre + "m"] = m
            cells[pre + "g1"] = g1
            cells[pre + "g2"] = g2
            cells[pre + "z"] = z
            cells[pre + "o"] = o
            cells[pre + "t"] = t
            cells[pre + "score"] = np.asarray(float(out['score']), np.float64)
            cells[pre + "path"] = np.asarray(out['alignment'].path, np.int32)
            cc += 1
    cells["n"] = np.asarray(cc)
    np.savez_compressed(os.path.join(HERE, "fill_cells.npz"), **cells)
    print("cell cases:", cc)


if __name__ == "__main__":
    main()
