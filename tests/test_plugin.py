"""The PRALINE plug-in boundary (B1 components, B2 batch manager) against the reference's own
components.  CPU part: interface mirror and error behaviour.  GPU part: identical outputs."""
import numpy as np
import pytest

import ref_praline as R
from praline_b200 import synth
from conftest import MODES

pytestmark = pytest.mark.skipif(not R.HAVE_PRALINE, reason="reference package (baseline/_ref) not present")

if R.HAVE_PRALINE:
    import praline
    import praline.component as pc
    from praline.core import Manager, ComponentError, DataError, Environment
    from praline.container import (Sequence, PlainTrack, ProfileTrack, ALPHABET_AA, ALPHABET_DNA, TRACK_ID_INPUT,
                                   TRACK_ID_PREPROFILE, MatchScoreModel, GapScoreModel)
    from praline_b200 import plugin


def _blosum():
    with praline.open_builtin('matrices/blosum62') as f:
        return praline.load_score_matrix(f, alphabet=ALPHABET_AA)


def _seq(name, idx, alpha=None):
    alpha = alpha or ALPHABET_AA
    return Sequence(name, [(TRACK_ID_INPUT, PlainTrack(None, alpha, raw_indices=np.asarray(idx)))])


# ---- CPU: the interface is the reference's ------------------------------------------------------
def test_components_mirror_reference_interface():
    for gpu, ref in ((plugin.GpuPairwiseAligner, pc.PairwiseAligner), (plugin.GpuRawPairwiseAligner, pc.RawPairwiseAligner)):
        assert gpu.tid == ref.tid
        assert set(gpu.inputs) == set(ref.inputs) and set(gpu.outputs) == set(ref.outputs)
        for k in ref.inputs:
            assert gpu.inputs[k].signature == ref.inputs[k].signature and gpu.inputs[k].optional == ref.inputs[k].optional
        assert gpu.options == ref.options and gpu.defaults == ref.defaults
    index = plugin.register(R.reference_index())
    assert index.resolve(pc.PairwiseAligner.tid) is plugin.GpuPairwiseAligner
    assert index.resolve(pc.RawPairwiseAligner.tid) is plugin.GpuRawPairwiseAligner


@pytest.mark.parametrize("case", ["track_sets", "multi_track", "gap_series", "alphabet", "matrix_dims"])
def test_same_errors_as_reference(case):
    sm = _blosum()
    a, b = _seq("a", [1, 2, 3]), _seq("b", [3, 2, 1])
    kw = dict(mode="global", sequence_one=a, sequence_two=b, track_id_sets_one=[[TRACK_ID_INPUT]],
              track_id_sets_two=[[TRACK_ID_INPUT]], score_matrices=[sm])
    env = {'gap_series': [-11.0, -1.0]}
    if case == "track_sets":
        kw['track_id_sets_two'] = [[TRACK_ID_INPUT], [TRACK_ID_INPUT]]
        kw['score_matrices'] = [sm, sm]
    elif case == "multi_track":
        kw['track_id_sets_one'] = [[TRACK_ID_INPUT, TRACK_ID_INPUT]]
    elif case == "gap_series":
        env = {'gap_series': [-11.0, -2.0, -1.0]}
    elif case == "alphabet":
        kw['sequence_two'] = _seq("b", [1, 2, 3], ALPHABET_DNA)
    elif case == "matrix_dims":
        class Fake(type(sm)):
            pass
        f = Fake.__new__(Fake)
        f.__dict__.update(sm.__dict__)
        f.matrix = np.zeros((27, 27, 27), np.float32)
        kw['score_matrices'] = [f]
    errs = []
    for index in (R.reference_index(), plugin.register(R.reference_index())):
        with pytest.raises((ComponentError, DataError)) as ei:
            R.run_task(Manager(index), pc.PairwiseAligner, env, **kw)
        errs.append((type(ei.value), str(ei.value)))
    assert errs[0] == errs[1]


def test_unknown_mode_is_component_error():
    sm = _blosum()
    a, b = _seq("a", [1, 2, 3]), _seq("b", [3, 2, 1])
    m = np.zeros((3, 3), np.float32)
    g = np.full((3, 2), -1, np.float32)
    for index in (R.reference_index(), plugin.register(R.reference_index())):
        with pytest.raises(ComponentError) as ei:
            R.run_task(Manager(index), pc.RawPairwiseAligner, {}, mode="bogus", sequence_one=a, sequence_two=b,
                       match_score_model=MatchScoreModel(a, b, m), gap_score_model_one=GapScoreModel(a, g),
                       gap_score_model_two=GapScoreModel(b, g))
        assert "unknown alignment mode: 'bogus'" in str(ei.value)


# ---- GPU: identical results through the component API ----------------------------------------------
def _same_alignment(got, want):
    assert type(got['score']) is float and got['score'] == want['score']
    gp, wp = got['alignment'].path, want['alignment'].path
    assert isinstance(gp, np.ndarray) == isinstance(wp, np.ndarray)
    assert np.array_equal(np.asarray(gp), np.asarray(wp))
    assert got['alignment'].items == want['alignment'].items


@pytest.mark.gpu
def test_pairwise_component_matches_reference():
    sm = _blosum()
    fam = synth.family(31, 6, 70)
    ref_mgr, gpu_mgr = Manager(R.reference_index()), Manager(plugin.register(R.reference_index()))
    for k, mode in enumerate(MODES):
        for gaps in ([-11.0, -1.0], [-8.0], [-2.5, -0.5]):
            a, b = _seq("a", fam[k]), _seq("b", fam[(k + 1) % 6])
            kw = dict(mode=mode, sequence_one=a, sequence_two=b, track_id_sets_one=[[TRACK_ID_INPUT]],
                      track_id_sets_two=[[TRACK_ID_INPUT]], score_matrices=[sm])
            zero = [(y, x) for y in range(3, 9) for x in range(2, 12)] if mode == "local" else None
            for z in (None, zero):
                if z is not None:
                    kw['zero_idxs'] = z
                want, wm = R.run_task(ref_mgr, pc.PairwiseAligner, {'gap_series': gaps}, **kw)
                got, gm = R.run_task(gpu_mgr, pc.PairwiseAligner, {'gap_series': gaps}, **kw)
                _same_alignment(got, want)
                assert [m.kind for m in gm][0] == "begin" and [m.kind for m in gm][-1] == "complete"


@pytest.mark.gpu
def test_profile_component_matches_reference():
    sm = _blosum()
    ref_mgr, gpu_mgr = Manager(R.reference_index()), Manager(plugin.register(R.reference_index()))
    for k, mode in enumerate(MODES):
        c1 = synth.count_profile(40 + k, 50 + 7 * k, 4 + k, 20, 27)
        c2 = synth.count_profile(60 + k, 64, 3, 20, 27)
        a = Sequence("p1", [(TRACK_ID_PREPROFILE, ProfileTrack(c1, ALPHABET_AA))])
        b = Sequence("p2", [(TRACK_ID_PREPROFILE, ProfileTrack(c2, ALPHABET_AA))])
        kw = dict(mode=mode, sequence_one=a, sequence_two=b, track_id_sets_one=[[TRACK_ID_PREPROFILE]],
                  track_id_sets_two=[[TRACK_ID_PREPROFILE]], score_matrices=[sm])
        want, _ = R.run_task(ref_mgr, pc.PairwiseAligner, {'gap_series': [-11.0, -1.0]}, **kw)
        got, _ = R.run_task(gpu_mgr, pc.PairwiseAligner, {'gap_series': [-11.0, -1.0]}, **kw)
        assert abs(got['score'] - want['score']) <= 1e-5 * max(1.0, abs(want['score']))
        _same_alignment(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["global", "semiglobal_both", "local"])
def test_long_profile_pair_through_the_component_matches_reference(mode):
    """A profile pair of 2^21 cells and more goes from GpuPairwiseAligner to GpuRawPairwiseAligner as OPERANDS and
    is aligned by one library call (pgpu_align_profile_long: score matrix beside the wavefront fill in the global and
    semiglobal modes, in sequence for local); same alignment and score as the reference's two components."""
    sm = _blosum()
    ref_mgr, gpu_mgr = Manager(R.reference_index()), Manager(plugin.register(R.reference_index()))
    c1 = synth.count_profile(140, 1500, 4, 20, 27)
    c2 = synth.count_profile(141, 1450, 3, 20, 27)
    a = Sequence("p1", [(TRACK_ID_PREPROFILE, ProfileTrack(c1, ALPHABET_AA))])
    b = Sequence("p2", [(TRACK_ID_PREPROFILE, ProfileTrack(c2, ALPHABET_AA))])
    kw = dict(mode=mode, sequence_one=a, sequence_two=b, track_id_sets_one=[[TRACK_ID_PREPROFILE]],
              track_id_sets_two=[[TRACK_ID_PREPROFILE]], score_matrices=[sm])
    want, _ = R.run_task(ref_mgr, pc.PairwiseAligner, {'gap_series': [-11.0, -1.0]}, **kw)
    got, _ = R.run_task(gpu_mgr, pc.PairwiseAligner, {'gap_series': [-11.0, -1.0]}, **kw)
    assert got['score'] == want['score']
    _same_alignment(got, want)


@pytest.mark.gpu
def test_raw_component_matches_reference():
    rng = np.random.default_rng(4)
    ref_mgr, gpu_mgr = Manager(R.reference_index()), Manager(plugin.register(R.reference_index()))
    a, b = _seq("a", rng.integers(0, 20, 45)), _seq("b", rng.integers(0, 20, 61))
    m = (rng.standard_normal((45, 61)) * 4).astype(np.float32)
    g1 = (-rng.random((45, 2)) * 5 - 0.1).astype(np.float32)
    g2 = (-rng.random((61, 2)) * 5 - 0.1).astype(np.float32)
    for mode in MODES:
        kw = dict(mode=mode, sequence_one=a, sequence_two=b, match_score_model=MatchScoreModel(a, b, m),
                  gap_score_model_one=GapScoreModel(a, g1), gap_score_model_two=GapScoreModel(b, g2),
                  zero_idxs=[(5, 5), (6, 6), (7, 7)])
        want, _ = R.run_task(ref_mgr, pc.RawPairwiseAligner, {}, **kw)
        got, _ = R.run_task(gpu_mgr, pc.RawPairwiseAligner, {}, **kw)
        _same_alignment(got, want)


@pytest.mark.gpu
def test_guide_tree_through_batch_manager():
    sm = _blosum()
    fam = synth.family(77, 14, 90)
    seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
    kw = dict(sequences=seqs, track_id_sets=[[TRACK_ID_INPUT]], score_matrices=[sm])
    for dist in ("global", "semiglobal", "semiglobal_auto"):
        env = {'gap_series': [-11.0, -1.0], 'linkage_method': 'average', 'dist_mode': dist,
               'aligner': pc.PairwiseAligner.tid}
        want, _ = R.run_task(Manager(R.reference_index()), pc.GuideTreeBuilder, env, **kw)
        # the reference's GuideTreeBuilder on the batching manager: one batched launch for its tasks
        mgr = plugin.GpuBatchManager(R.reference_index(), gpu_tree=False)
        got, msgs = R.run_task(mgr, pc.GuideTreeBuilder, env, **kw)
        assert got['guide_tree'].merge_orders == want['guide_tree'].merge_orders
        assert mgr.batched_requests == 14 * 13 // 2
        assert sum(1 for m in msgs if m.kind == "complete") >= 14 * 13 // 2
        # the GPU GuideTreeBuilder (all-vs-all launch + clustering kernel) under the same type id
        mgr = plugin.GpuBatchManager(R.reference_index())
        got, msgs = R.run_task(mgr, pc.GuideTreeBuilder, env, **kw)
        assert got['guide_tree'].merge_orders == want['guide_tree'].merge_orders
        assert all(isinstance(a, int) and isinstance(b, int) for a, b in got['guide_tree'].merge_orders)
        assert mgr.batched_requests == (0 if dist != "semiglobal_auto" else 14 * 13 // 2)


@pytest.mark.gpu
@pytest.mark.parametrize("linkage", ["single", "complete", "average"])
def test_gpu_guide_tree_on_profile_tracks(linkage):
    """The tree stage of the MSA workflow runs on preprofile tracks (workflow.py:164-181): f32
    profile x profile scores, clustered on the device, same merge order as the reference."""
    sm = _blosum()
    rng = np.random.default_rng(5)
    seqs = []
    for i in range(11):
        L = int(rng.integers(40, 70))
        counts = np.zeros((L, ALPHABET_AA.size), np.int64)
        for _ in range(int(rng.integers(1, 6))):
            counts[np.arange(L), rng.integers(0, 20, L)] += 1
        seqs.append(Sequence("p%d" % i, [(TRACK_ID_PREPROFILE, ProfileTrack(counts, ALPHABET_AA))]))
    kw = dict(sequences=seqs, track_id_sets=[[TRACK_ID_PREPROFILE]], score_matrices=[sm])
    env = {'gap_series': [-11.0, -1.0], 'linkage_method': linkage, 'dist_mode': 'global',
           'aligner': pc.PairwiseAligner.tid}
    want, _ = R.run_task(Manager(R.reference_index()), pc.GuideTreeBuilder, env, **kw)
    got, _ = R.run_task(plugin.GpuBatchManager(R.reference_index()), pc.GuideTreeBuilder, env, **kw)
    assert got['guide_tree'].merge_orders == want['guide_tree'].merge_orders


def test_guide_tree_component_mirrors_reference():
    ref, gpu = pc.GuideTreeBuilder, plugin.GpuGuideTreeBuilder
    assert gpu.tid == ref.tid
    assert set(gpu.inputs) == set(ref.inputs) and set(gpu.outputs) == set(ref.outputs)
    assert gpu.options == ref.options
    assert {k: v for k, v in gpu.defaults.items() if k != 'aligner_env'} == \
           {k: v for k, v in ref.defaults.items() if k != 'aligner_env'}


@pytest.mark.gpu
def test_lazy_alignment_paths_from_batch():
    sm = _blosum()
    fam = synth.family(78, 6, 50)
    seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
    from praline.core import Execution
    for mode in ("global", "semiglobal_both", "local"):
        outs = []
        for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
            ex = Execution(mgr, R.ROOT_TAG)
            for i in range(6):
                for j in range(6):
                    if i != j:
                        t = ex.add_task(pc.PairwiseAligner)
                        t.environment(Environment(keys={'gap_series': [-11.0, -1.0]}))
                        t.inputs(mode=mode, sequence_one=seqs[i], sequence_two=seqs[j],
                                 track_id_sets_one=[[TRACK_ID_INPUT]], track_id_sets_two=[[TRACK_ID_INPUT]],
                                 score_matrices=[sm])
            for _ in ex.run():
                pass
            outs.append(ex.outputs)
        for want, got in zip(*outs):
            _same_alignment(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("preprofile,msa", [("global", "tree"), ("dummy", "tree"), ("global", "ad_hoc"),
                                            ("local", "tree")])
def test_msa_workflow_identical_to_reference(preprofile, msa):
    sm = _blosum()
    seqs = praline.load_sequence_fasta(R.ROOT + "/tests/golden/BBA0184.tfa", ALPHABET_AA)
    want = R.workflow_fasta(Manager(R.reference_index()), seqs, sm, preprofile, msa)
    seqs = praline.load_sequence_fasta(R.ROOT + "/tests/golden/BBA0184.tfa", ALPHABET_AA)
    got = R.workflow_fasta(plugin.GpuBatchManager(R.reference_index()), seqs, sm, preprofile, msa)
    assert got == want
    if (preprofile, msa) == ("global", "tree"):
        assert got == open(R.ROOT + "/tests/golden/BBA0184.aln").read()     # the reference's golden file
    fam = synth.family(5, 12, 60)
    mk = lambda: [_seq("s%d" % i, s) for i, s in enumerate(fam)]
    want = R.workflow_fasta(Manager(R.reference_index()), mk(), sm, preprofile, msa)
    got = R.workflow_fasta(plugin.GpuBatchManager(R.reference_index()), mk(), sm, preprofile, msa)
    assert got == want


@pytest.mark.gpu
def test_long_sequences_fall_back_to_general_kernels():
    """Sequences beyond the inter-task limit (1024) still align identically through the plug-in."""
    sm = _blosum()
    rng = np.random.default_rng(12)
    a, b = _seq("a", rng.integers(0, 20, 1300)), _seq("b", rng.integers(0, 20, 1190))
    kw = dict(mode="global", sequence_one=a, sequence_two=b, track_id_sets_one=[[TRACK_ID_INPUT]],
              track_id_sets_two=[[TRACK_ID_INPUT]], score_matrices=[sm])
    want, _ = R.run_task(Manager(R.reference_index()), pc.PairwiseAligner, {'gap_series': [-11.0, -1.0]}, **kw)
    got, _ = R.run_task(plugin.GpuBatchManager(R.reference_index()), pc.PairwiseAligner, {'gap_series': [-11.0, -1.0]}, **kw)
    _same_alignment(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("threshold", [None, 250.0])
def test_bulk_master_slave_matches_reference(threshold):
    """N GlobalMasterSlaveAligner requests in one Execution (workflow.py:139-161) take the bulk
    path: the lazily built alignments and the device count tables (ProfileBuilder) equal the
    reference's, with and without a score threshold (preprofile.py:145)."""
    from praline.core import Execution
    sm = _blosum()
    fam = synth.family(31, 9, 70)
    outs, profs = [], []
    for mk_mgr in (lambda: Manager(R.reference_index()), lambda: plugin.GpuBatchManager(R.reference_index())):
        mgr = mk_mgr()
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
        ex = Execution(mgr, R.ROOT_TAG)
        for i, master in enumerate(seqs):
            t = ex.add_task(pc.GlobalMasterSlaveAligner)
            t.environment(Environment(keys={'gap_series': [-11.0, -1.0], 'score_threshold': threshold,
                                            'aligner': pc.PairwiseAligner.tid}))
            t.inputs(master_sequence=master, slave_sequences=[s for j, s in enumerate(seqs) if j != i],
                     track_id_sets=[[TRACK_ID_INPUT]], score_matrices=[sm])
        for _ in ex.run():
            pass
        alns = [o['alignment'] for o in ex.outputs]
        # ProfileBuilder on every alignment (the device count tables on the GPU manager) ...
        ex2 = Execution(mgr, R.ROOT_TAG)
        for a in alns:
            t = ex2.add_task(pc.ProfileBuilder)
            t.environment(Environment(keys={}))
            t.inputs(alignment=a, track_id=TRACK_ID_INPUT)
        for _ in ex2.run():
            pass
        profs.append([o['profile_track'].counts for o in ex2.outputs])
        # ... and the alignments themselves (lazy on the GPU manager)
        outs.append([(np.asarray(a.path), [it.name for it in a.items]) for a in alns])
        if isinstance(mgr, plugin.GpuBatchManager):
            assert mgr.batched_requests == 9 * 8
    for (p0, n0), (p1, n1) in zip(*outs):
        assert n0 == n1 and np.array_equal(p0, p1)
    for c0, c1 in zip(*profs):
        assert np.array_equal(c0, c1)


@pytest.mark.gpu
@pytest.mark.parametrize("threshold,iterations", [(None, 2), (60.0, 3), (None, 1)])
def test_bulk_local_master_slave_matches_reference(threshold, iterations):
    """N LocalMasterSlaveAligner requests in one Execution (workflow.py:139-161, preprofile_mode
    'local') take the bulk path: Waterman-Eggert iterations with their box masks on the device;
    the lazily built alignments and the device count tables equal the reference's."""
    from praline.core import Execution
    sm = _blosum()
    fam = synth.family(33, 7, 60) + [np.random.default_rng(3).integers(0, 20, 21)]
    outs, profs = [], []
    for mk_mgr in (lambda: Manager(R.reference_index()), lambda: plugin.GpuBatchManager(R.reference_index())):
        mgr = mk_mgr()
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
        ex = Execution(mgr, R.ROOT_TAG)
        for i, master in enumerate(seqs):
            t = ex.add_task(pc.LocalMasterSlaveAligner)
            t.environment(Environment(keys={'gap_series': [-11.0, -1.0], 'score_threshold': threshold,
                                            'aligner': pc.PairwiseAligner.tid,
                                            'waterman_eggert_iterations': iterations}))
            t.inputs(master_sequence=master, slave_sequences=[s for j, s in enumerate(seqs) if j != i],
                     track_id_sets=[[TRACK_ID_INPUT]], score_matrices=[sm])
        for _ in ex.run():
            pass
        alns = [o['alignment'] for o in ex.outputs]
        ex2 = Execution(mgr, R.ROOT_TAG)
        for a in alns:
            t = ex2.add_task(pc.ProfileBuilder)
            t.environment(Environment(keys={}))
            t.inputs(alignment=a, track_id=TRACK_ID_INPUT)
        for _ in ex2.run():
            pass
        profs.append([o['profile_track'].counts for o in ex2.outputs])
        outs.append([(np.asarray(a.path), [it.name for it in a.items]) for a in alns])
        if isinstance(mgr, plugin.GpuBatchManager):
            assert mgr.batched_requests == 8 * 7 * iterations
    for (p0, n0), (p1, n1) in zip(*outs):
        assert n0 == n1 and np.array_equal(p0, p1)
    for c0, c1 in zip(*profs):
        assert np.array_equal(c0, c1)


def test_vectorised_merges_equal_reference_loops():
    """merge_profile_counts / merge_alignment_paths against ProfileTrack.merge
    (container/sequence.py:205-239) and Alignment.merge (container/align.py:30-61) on random
    monotone paths."""
    from praline.container import Alignment
    rng = np.random.default_rng(1)
    for trial in range(100):
        L1, L2 = (int(v) for v in rng.integers(1, 30, 2))
        y = x = 0
        rows = [(0, 0)]
        while (y, x) != (L1, L2):
            opts = [(1, 1)] * (y < L1 and x < L2) + [(1, 0)] * (y < L1) + [(0, 1)] * (x < L2)
            dy, dx = opts[int(rng.integers(len(opts)))]
            y, x = y + dy, x + dx
            rows.append((y, x))
        path = np.array(rows)
        t1 = ProfileTrack(rng.integers(0, 5, (L1, 27)), ALPHABET_AA)
        t2 = ProfileTrack(rng.integers(0, 5, (L2, 27)), ALPHABET_AA)
        want = t1.merge(t2, path).counts
        got = ProfileTrack(plugin.merge_profile_counts(t1.counts, t2.counts, path), ALPHABET_AA).counts
        assert np.array_equal(got, want)
        n1, n2 = (int(v) for v in rng.integers(1, 4, 2))
        p1, p2 = rng.integers(-1, 9, (L1 + 1, n1)), rng.integers(-1, 9, (L2 + 1, n2))
        want = Alignment(list(range(n1)), p1).merge(Alignment(list(range(n2)), p2), path).path
        got = plugin.merge_alignment_paths(p1, p2, path)
        assert np.array_equal(got, want) and got.dtype == want.dtype


def test_tree_msa_component_mirrors_reference():
    ref, gpu = pc.TreeMultipleSequenceAligner, plugin.GpuTreeMultipleSequenceAligner
    assert gpu.tid == ref.tid
    assert set(gpu.inputs) == set(ref.inputs) and set(gpu.outputs) == set(ref.outputs)
    assert gpu.options == ref.options
    assert {k: v for k, v in gpu.defaults.items() if k != 'aligner_env'} == \
           {k: v for k, v in ref.defaults.items() if k != 'aligner_env'}


@pytest.mark.gpu
@pytest.mark.parametrize("merge_mode", ["global", "semiglobal", "semiglobal_auto"])
def test_tree_msa_component_matches_reference(merge_mode):
    """TreeMultipleSequenceAligner on a fixed guide tree: the GPU component (vectorised merges,
    K1 + K3 alignments) returns the reference's alignment path and item order."""
    sm = _blosum()
    fam = synth.family(61, 10, 55)
    from praline.container import SequenceTree
    tree_orders = [(0, 3), (1, 2), (4, 9), (0, 1), (5, 6), (7, 8), (0, 4), (5, 7), (0, 5)]
    outs = []
    for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
        env = {'gap_series': [-11.0, -1.0], 'merge_mode': merge_mode, 'aligner': pc.PairwiseAligner.tid}
        out, _ = R.run_task(mgr, pc.TreeMultipleSequenceAligner, env, sequences=seqs,
                            guide_tree=SequenceTree(seqs, tree_orders), track_id_sets=[[TRACK_ID_INPUT]],
                            score_matrices=[sm])
        outs.append(out['alignment'])
    assert [s.name for s in outs[0].items] == [s.name for s in outs[1].items]
    assert np.array_equal(np.asarray(outs[0].path), np.asarray(outs[1].path))


@pytest.mark.gpu
@pytest.mark.parametrize("merge_mode", ["global", "semiglobal"])
def test_tree_msa_device_resident_levels_equal_per_merge_path(merge_mode, monkeypatch):
    """The device-resident, level-batched merge path (Engine.merge_level: count tables never leave the device,
    one launch group per guide-tree level) gives the alignment of the per-merge path (every merge through the
    manager as a PairwiseAligner task) on preprofile tracks of 36 sequences and a guide tree with levels of
    several independent merges; a 12-sequence subset is also checked against the reference's own component."""
    from praline.container import SequenceTree, ProfileTrack, ALPHABET_AA
    sm = _blosum()
    rng = np.random.default_rng(5)
    fam = synth.family(71, 36, 70)

    def mk(k):
        seqs = []
        for i, s in enumerate(fam[:k]):
            q = _seq("s%d" % i, s)
            counts = np.zeros((len(s), 27), np.int64)          # a preprofile-like count track: own residue + noise
            counts[np.arange(len(s)), s] = 6
            counts[np.arange(len(s)), rng.integers(0, 20, len(s))] += rng.integers(0, 4, len(s))
            q.add_track("prof", ProfileTrack(counts, ALPHABET_AA))
            seqs.append(q)
        return seqs

    def orders(k):          # pairs first (one wide level), then a chain over the survivors
        o = [(i, i + 1) for i in range(0, k - 1, 2)]
        alive = list(range(0, k, 2))
        while len(alive) > 1:
            nxt = []
            for a in range(0, len(alive) - 1, 2):
                o.append((alive[a], alive[a + 1]))
                nxt.append(alive[a])
            if len(alive) % 2:
                nxt.append(alive[-1])
            alive = nxt
        return o

    def run(mgr, k, track):
        seqs = mk(k)
        env = {'gap_series': [-11.0, -1.0], 'merge_mode': merge_mode, 'aligner': pc.PairwiseAligner.tid}
        out, _ = R.run_task(mgr, pc.TreeMultipleSequenceAligner, env, sequences=seqs,
                            guide_tree=SequenceTree(seqs, orders(k)), track_id_sets=[[track]], score_matrices=[sm])
        return [s.name for s in out['alignment'].items], np.asarray(out['alignment'].path)

    for track in ("prof", TRACK_ID_INPUT):
        rng = np.random.default_rng(5)
        fast = run(plugin.GpuBatchManager(R.reference_index()), 36, track)
        monkeypatch.setenv("PGPU_NO_DEVICE_MERGE", "1")
        rng = np.random.default_rng(5)
        slow = run(plugin.GpuBatchManager(R.reference_index()), 36, track)
        monkeypatch.delenv("PGPU_NO_DEVICE_MERGE")
        assert fast[0] == slow[0] and np.array_equal(fast[1], slow[1]), track
    rng = np.random.default_rng(5)
    got = run(plugin.GpuBatchManager(R.reference_index()), 12, "prof")
    rng = np.random.default_rng(5)
    want = run(Manager(R.reference_index()), 12, "prof")
    assert got[0] == want[0] and np.array_equal(got[1], want[1])


@pytest.mark.gpu
def test_pairwise_aligner_message_sequence_matches_reference():
    """SURVEY 4 message sequence: a PairwiseAligner task yields Begin(self), Begin(RawPairwiseAligner),
    [Log(Raw)], Complete(Raw), [Log(self)], Complete(self) -- the same kinds, tag classes and parent links from the
    GPU components on the single-request fast path, the general path with debug = 1 (both log bundles) and the
    batched execute_many path."""
    from praline.core import Execution, Environment
    sm = _blosum()
    fam = synth.family(13, 4, 40)

    def shape(msgs):
        out = []
        tags = {}
        for m in msgs:
            cls = m.tag.split("#")[0]
            tags.setdefault(m.tag, len(tags))
            parent = getattr(m, "parent_tag", None)
            out.append((m.kind, cls, tags[m.tag], tags.get(parent) if parent in tags else (parent.split("#")[0] if parent else None)))
        return out

    for debug in (0, 1):
        got = []
        for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
            seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
            env = {'gap_series': [-11.0, -1.0], 'debug': debug}
            _, msgs = R.run_task(mgr, pc.PairwiseAligner, env, mode="global", sequence_one=seqs[0], sequence_two=seqs[1],
                                 track_id_sets_one=[[TRACK_ID_INPUT]], track_id_sets_two=[[TRACK_ID_INPUT]],
                                 score_matrices=[sm])
            got.append(shape(msgs))
        assert got[0] == got[1], debug
    got = []
    for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):     # three tasks: batched
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
        ex = Execution(mgr, R.ROOT_TAG)
        for a, b in ((0, 1), (0, 2), (1, 3)):
            t = ex.add_task(pc.PairwiseAligner)
            t.environment(Environment(keys={'gap_series': [-11.0, -1.0], 'debug': 0}))
            t.inputs(mode="global", sequence_one=seqs[a], sequence_two=seqs[b], track_id_sets_one=[[TRACK_ID_INPUT]],
                     track_id_sets_two=[[TRACK_ID_INPUT]], score_matrices=[sm])
        got.append(shape([m for m in ex.run()]))
    assert got[0] == got[1]


@pytest.mark.gpu
def test_local_with_open_weaker_than_extend_matches_reference():
    """gap_series [-1, -5]: the border cell o[0,0,1] = open - extend = 4 is positive and takes part in the reference's
    argmax over the whole o array (component/align.py:371, :401-403).  Such requests must not ride the batched local
    reduction (interior cells only); single requests and execute_many batches both have to give the reference's
    score and path -- including pairs where nothing in the interior beats the border."""
    from praline.core import Execution, Environment
    sm = _blosum()
    G, W = 7, 17            # glycine / tryptophan: S[G][W] = -2, no positive cell
    rng = np.random.default_rng(4)
    seqs_idx = [np.full(6, G), np.full(5, W), rng.integers(0, 20, 30), rng.integers(0, 20, 28)]
    outs = []
    for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(seqs_idx)]
        ex = Execution(mgr, R.ROOT_TAG)
        for a, b in ((0, 1), (2, 3), (0, 2), (1, 3)):
            t = ex.add_task(pc.PairwiseAligner)
            t.environment(Environment(keys={'gap_series': [-1.0, -5.0], 'debug': 0}))
            t.inputs(mode="local", sequence_one=seqs[a], sequence_two=seqs[b], track_id_sets_one=[[TRACK_ID_INPUT]],
                     track_id_sets_two=[[TRACK_ID_INPUT]], score_matrices=[sm])
        for _ in ex.run():
            pass
        outs.append([(o['score'], [tuple(p) for p in o['alignment'].path]) for o in ex.outputs])
    assert outs[0] == outs[1]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3])
def test_gpu_guide_tree_tiny_inputs(n):
    """One, two and three sequences through the GPU GuideTreeBuilder and the tree MSA."""
    sm = _blosum()
    fam = synth.family(91, 3, 40)[:n]
    outs = []
    for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
        env = {'gap_series': [-11.0, -1.0], 'linkage_method': 'average', 'dist_mode': 'global',
               'aligner': pc.PairwiseAligner.tid, 'merge_mode': 'global'}
        tree, _ = R.run_task(mgr, pc.GuideTreeBuilder, env, sequences=seqs, track_id_sets=[[TRACK_ID_INPUT]],
                             score_matrices=[sm])
        msa, _ = R.run_task(mgr, pc.TreeMultipleSequenceAligner, env, sequences=seqs, guide_tree=tree['guide_tree'],
                            track_id_sets=[[TRACK_ID_INPUT]], score_matrices=[sm])
        outs.append((tree['guide_tree'].merge_orders, np.asarray(msa['alignment'].path)))
    assert list(map(tuple, outs[0][0])) == list(map(tuple, outs[1][0]))
    assert np.array_equal(outs[0][1], outs[1][1])


def test_adhoc_msa_component_mirrors_reference():
    ref, gpu = pc.AdHocMultipleSequenceAligner, plugin.GpuAdHocMultipleSequenceAligner
    assert gpu.tid == ref.tid
    assert set(gpu.inputs) == set(ref.inputs) and set(gpu.outputs) == set(ref.outputs)
    assert gpu.options == ref.options
    assert {k: v for k, v in gpu.defaults.items() if k != 'aligner_env'} == \
           {k: v for k, v in ref.defaults.items() if k != 'aligner_env'}


@pytest.mark.gpu
@pytest.mark.parametrize("dist_mode,merge_mode", [("global", "semiglobal"), ("semiglobal", "global"),
                                                  ("global", "semiglobal_auto")])
def test_adhoc_msa_component_matches_reference(dist_mode, merge_mode):
    """AdHocMultipleSequenceAligner (the CLI's default MSA mode): the GPU component (score matrix
    cache on ids, batched rounds, vectorised merges) returns the reference's alignment."""
    sm = _blosum()
    fam = synth.family(62, 13, 48)
    outs = []
    for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
        seqs = [_seq("s%d" % i, s) for i, s in enumerate(fam)]
        env = {'gap_series': [-11.0, -1.0], 'merge_mode': merge_mode, 'dist_mode': dist_mode,
               'aligner': pc.PairwiseAligner.tid}
        out, _ = R.run_task(mgr, pc.AdHocMultipleSequenceAligner, env, sequences=seqs,
                            track_id_sets=[[TRACK_ID_INPUT]], score_matrices=[sm])
        outs.append(out['alignment'])
    assert [s.name for s in outs[0].items] == [s.name for s in outs[1].items]
    assert np.array_equal(np.asarray(outs[0].path), np.asarray(outs[1].path))


@pytest.mark.gpu
def test_adhoc_msa_on_profile_tracks_matches_reference():
    """Ad-hoc MSA on preprofile tracks (what --preprofile-global feeds it): f32 profile scores in the
    reference's evaluation order decide every merge."""
    sm = _blosum()
    rng = np.random.default_rng(8)
    def mk():
        r = np.random.default_rng(8)
        seqs = []
        for i in range(9):
            L = int(r.integers(30, 50))
            counts = np.zeros((L, ALPHABET_AA.size), np.int64)
            for _ in range(int(r.integers(1, 5))):
                counts[np.arange(L), r.integers(0, 20, L)] += 1
            seqs.append(Sequence("p%d" % i, [(TRACK_ID_PREPROFILE, ProfileTrack(counts, ALPHABET_AA))]))
        return seqs
    outs = []
    for mgr in (Manager(R.reference_index()), plugin.GpuBatchManager(R.reference_index())):
        env = {'gap_series': [-11.0, -1.0], 'merge_mode': 'semiglobal', 'dist_mode': 'global',
               'aligner': pc.PairwiseAligner.tid}
        out, _ = R.run_task(mgr, pc.AdHocMultipleSequenceAligner, env, sequences=mk(),
                            track_id_sets=[[TRACK_ID_PREPROFILE]], score_matrices=[sm])
        outs.append(out['alignment'])
    assert [s.name for s in outs[0].items] == [s.name for s in outs[1].items]
    assert np.array_equal(np.asarray(outs[0].path), np.asarray(outs[1].path))
