"""PRALINE plug-in: GPU pairwise aligner components and a batching manager.

Drop-in boundary.  Needs the `praline` package (the reference) importable; it
is not imported by praline_b200 itself.

B1  GpuPairwiseAligner / GpuRawPairwiseAligner carry the SAME type ids, ports, options and
    defaults as the reference's PairwiseAligner / RawPairwiseAligner
    (praline/component/align.py:75-86, 289-300), raise the same ComponentError / DataError
    for the same conditions (:114-189, :319-341) and return the same container types
    (Alignment with a list-of-tuples path for global/local, an int ndarray for semiglobal,
    :401-433; score a Python float).  Registering them in a TypeIndex replaces the CPU
    aligners for every caller that resolves env['aligner'] (preprofile.py:129, tree.py:117,
    msa.py:175,421,529).
B2  GpuBatchManager(Manager).execute_many (praline/core/manager.py:154-170) sees the whole
    list of requests of one Execution (execution.py:164-175).  PairwiseAligner requests on
    sequence tracks are aligned by ONE batched kernel launch per group; everything else goes
    through the stock Manager path.  N GlobalMasterSlaveAligner requests
    (workflow.py:139-161) are flattened into one N(N-1) ordered-pair batch.

All arithmetic runs in libpraline_b200.so; without a GPU these components raise.
"""
from __future__ import division, absolute_import, print_function

import numpy as np

from praline.core import (Component, Port, Environment, Execution, Manager, T, BeginMessage, CompleteMessage,
                          ProgressMessage, LogMessage, ComponentError, DataError, LogBundle, ROOT_LOG_NAME,
                          path_to_url)
from praline.container import (Sequence, Alignment, ScoreMatrix, PlainTrack, ProfileTrack, MatchScoreModel,
                               GapScoreModel, SequenceTree)
from praline.util import compress_path, extend_path_local, auto_align_mode
from praline.container import TRACK_ID_INPUT

from . import _lib
from .engine import get_engine, MODES

PAIRWISE_TID = "praline.component.PairwiseAligner"
RAW_TID = "praline.component.RawPairwiseAligner"
GLOBAL_MS_TID = "praline.component.GlobalMasterSlaveAligner"
LOCAL_MS_TID = "praline.component.LocalMasterSlaveAligner"
PROFILE_BUILDER_TID = "praline.component.ProfileBuilder"
GUIDE_TREE_TID = "praline.component.GuideTreeBuilder"
TREE_MSA_TID = "praline.component.TreeMultipleSequenceAligner"
ADHOC_MSA_TID = "praline.component.AdHocMultipleSequenceAligner"


def _path_container(mode, path):
    """Reference container types: list of 2-tuples for global/local (util/align.py:183), int
    ndarray for the semiglobal modes (component/align.py:425-426)."""
    if mode in ("global", "local"):
        return list(map(tuple, np.asarray(path).tolist()))
    return np.asarray(path, dtype=int)


class LazyAlignment(Alignment):
    """An Alignment whose path is produced on first access.  GuideTreeBuilder only reads the
    score of its N(N-1)/2 alignments (tree.py:142-145), so the batched path is score-only
    until somebody asks for a path; the first access traces the whole batch at once."""
    tid = Alignment.tid

    def __init__(self, items, resolver, key):
        self.items = items
        self._resolver = resolver
        self._key = key
        self._path = None

    @property
    def path(self):
        if self._path is None:
            self._path = self._resolver(self._key)
            self._resolver = None
        return self._path

    @path.setter
    def path(self, value):
        self._path = value

    def __reduce__(self):
        return (Alignment, (self.items, self.path))


class _PreprofileBatch(object):
    """All master-slave pairs of one execute_many call that share a (matrix, gaps, threshold)
    class.  The count tables of every master are produced by ONE device pass on first use."""

    def __init__(self, group, threshold, iterations=None):
        self.group, self.threshold = group, threshold
        self.iterations = iterations      # None: global master-slave; n: local with n Waterman-Eggert iterations
        self.masters, self.slaves = [], []
        self.result = None
        self._local = None

    def add(self, master_idx, slave_idx):
        self.masters.append(master_idx)
        self.slaves.append(slave_idx)

    def add_many(self, master_idx, slave_idxs):
        self.masters.extend([master_idx] * len(slave_idxs))
        self.slaves.extend(slave_idxs)

    def counts(self, master_idx):
        if self.result is None:
            g = self.group
            if self.iterations is None:
                self.result = get_engine().preprofile_counts(g.batch, self.masters, self.slaves, g.S, g.gaps,
                                                             self.threshold)
            else:
                self.result = get_engine().local_preprofile_counts(g.batch, self.masters, self.slaves, g.S, g.gaps,
                                                                   iterations=self.iterations, threshold=self.threshold)
        cnt, where, _ = self.result
        off, length = where[int(master_idx)]
        A = self.group.S.shape[0]
        return cnt[off:off + length * A].reshape(length, A).copy()


    def local_alignments(self):
        """(scores [iterations x pairs], paths per iteration) of the local batch, traced on first use."""
        if self._local is None:
            g = self.group
            scores, paths, _ = get_engine().local_pairs(g.batch, self.masters, self.slaves, g.S, g.gaps,
                                                        iterations=self.iterations, want_paths=True)
            self._local = (scores, paths)
        return self._local


class _StagePreprofile(object):
    """The whole preprofile stage of a workflow (workflow.py:66-73: every sequence is a master, its
    slaves are all the others in input order): no pair list is ever built on the host -- the count
    tables come from Engine.preprofile_stage (masters in chunks, nothing read back in between), and a
    master's alignments are traced on demand for the rare caller that wants the alignment itself."""
    complete = True

    def __init__(self, group, threshold, iterations=None):
        self.group, self.threshold, self.iterations = group, threshold, iterations
        self.result = None

    def counts(self, master_idx):
        if self.result is None:
            g = self.group
            cnt, where, _ = get_engine().preprofile_stage(g.batch, g.S, g.gaps, threshold=self.threshold,
                                                          mode="global" if self.iterations is None else "local",
                                                          iterations=self.iterations or 1)
            self.result = (cnt.cpu().numpy(), where)
        cnt, where = self.result
        off, length = where[int(master_idx)]
        A = self.group.S.shape[0]
        return cnt[off:off + length * A].reshape(length, A).astype(np.int64)

    def alignments_of(self, master_idx):
        """(scores [iterations x n-1], paths [iterations][n-1]) of one master against all others."""
        g = self.group
        n = len(g.seqs)
        sidx = np.r_[0:master_idx, master_idx + 1:n]
        midx = np.full(n - 1, master_idx, np.int64)
        if self.iterations is None:
            scores, paths = get_engine().align_pairs(g.batch, midx, sidx, g.S, g.gaps, mode="global", want_paths=True,
                                                     resident="one")
            return scores[None, :], [paths]
        scores, paths, _ = get_engine().local_pairs(g.batch, midx, sidx, g.S, g.gaps, iterations=self.iterations,
                                                    want_paths=True)
        return scores, paths


class _RangePicks(object):
    """(group, pair number) of every slave of one master: pairs k0 .. k0+n-1 of the group."""

    def __init__(self, group, k0, n):
        self.group, self.k0, self.n = group, k0, n

    def __iter__(self):
        return iter((self.group, k) for k in range(self.k0, self.k0 + self.n))

    def __len__(self):
        return self.n


class LazyMasterSlaveAlignment(Alignment):
    """The master-slave alignment of GlobalMasterSlaveAligner (preprofile.py:67-156), built on
    first access.  The workflow only hands it to ProfileBuilder (workflow.py:211-224), whose
    count table GpuBatchManager takes from the device (`gpu_counts`) without ever materialising
    the paths; anybody else who touches .items / .path gets the reference's host-side merge."""
    tid = Alignment.tid

    def __init__(self, master, slaves, picks, threshold, track_id, batch, master_idx):
        self._master, self._slaves, self._picks = master, slaves, picks
        self._threshold, self._track_id = threshold, track_id
        self._batch, self._master_idx = batch, master_idx
        self._built = None

    def _build(self):
        if self._built is None:
            master = self._master
            path = np.arange(len(master) + 1).reshape(len(master) + 1, 1)
            alignment = Alignment([master], path)
            if getattr(self._batch, "complete", False):     # preprofile.py:127-154 / :227-265, this master only
                scores, paths = self._batch.alignments_of(self._master_idx)
                local = self._batch.iterations is not None
                for k, slave in enumerate(self._slaves):
                    for n in range(len(paths)):
                        if self._threshold is None or float(scores[n][k]) >= self._threshold:
                            p = compress_path(np.array(paths[n][k], dtype=int), 0)
                            if local:
                                p = extend_path_local(p, len(master), 0)
                            merge_path = np.arange(len(slave) + 1).reshape(len(slave) + 1, 1)
                            alignment = alignment.merge(Alignment([slave], merge_path), p)
                self._built = alignment
                return self._built
            if self._batch is not None and self._batch.iterations is not None:    # preprofile.py:227-265
                scores, paths = self._batch.local_alignments()
                for slave, (g, k) in zip(self._slaves, self._picks):
                    for n in range(self._batch.iterations):
                        if self._threshold is None or float(scores[n][k]) >= self._threshold:
                            p = compress_path(np.array(paths[n][k], dtype=int), 0)
                            p = extend_path_local(p, len(master), 0)
                            merge_path = np.arange(len(slave) + 1).reshape(len(slave) + 1, 1)
                            alignment = alignment.merge(Alignment([slave], merge_path), p)
                self._built = alignment
                return self._built
            for slave, (g, k) in zip(self._slaves, self._picks):     # preprofile.py:144-152
                score = float(g.scores[k])
                if self._threshold is None or score >= self._threshold:
                    p = compress_path(np.array(g.path(k)), 0)
                    merge_path = np.arange(len(slave) + 1).reshape(len(slave) + 1, 1)
                    alignment = alignment.merge(Alignment([slave], merge_path), p)
            self._built = alignment
        return self._built

    @property
    def items(self):
        return self._build().items

    @property
    def path(self):
        return self._build().path

    def gpu_counts(self, track_id):
        """Count table of ProfileBuilder (profile.py:56) or None when the device path does not apply."""
        if track_id != self._track_id or self._built is not None:
            return None
        return self._batch.counts(self._master_idx)

    def __reduce__(self):
        return (Alignment, (self.items, self.path))


def _seq_like(track):
    """Index sequence of a track if it is a plain track or a one-symbol-per-column profile
    (then P.S.P^T is exactly S[a][b], cext.c:84-89), else None."""
    if track.tid == PlainTrack.tid:
        return np.asarray(track.values)
    if track.tid == ProfileTrack.tid:
        counts = track.counts
        if counts.ndim == 2 and counts.shape[0] > 0 and ((counts != 0).sum(axis=1) == 1).all() and (counts >= 0).all():
            return np.argmax(counts != 0, axis=1).astype(np.int32)
    return None


def _prepare(sequence_one, sequence_two, track_id_sets_one, track_id_sets_two, score_matrices, gap_series):
    """The checks and array preparation of PairwiseAligner.execute (component/align.py:114-189),
    same exceptions, same messages.  Returns (sets, gaps) with sets = [(track1, track2, S)]."""
    if len(track_id_sets_one) != len(track_id_sets_two):
        raise ComponentError("should have an identical number of track id sets"
                             "for both sequences")
    sets = []
    for n, (track_ids_one, track_ids_two) in enumerate(zip(track_id_sets_one, track_id_sets_two)):
        score_matrix = score_matrices[n]
        if len(track_ids_one) != 1 or len(track_ids_two) != 1:
            raise ComponentError("the fast aligner only supports single-track"
                                 " alignments at the moment")
        total = len(track_ids_one) + len(track_ids_two)
        if score_matrix.matrix.ndim != total:
            s = "the score matrix must consist of as many dimensions as" \
                "there are tracks to be aligned ({0}), but it contains " \
                "{1}"
            raise ComponentError(s.format(total, score_matrix.matrix.ndim))
        track_one = sequence_one.get_track(track_ids_one[0])
        track_two = sequence_two.get_track(track_ids_two[0])
        if score_matrix.alphabets[0].aid != track_one.alphabet.aid:
            s = "track {0} for sequence one has alphabet '{1}' but " \
                "the corresponding dimension in the score matrix " \
                "has alphabet '{2}'"
            raise DataError(s.format(0, track_one.alphabet.aid, score_matrix.alphabets[0].aid))
        if score_matrix.alphabets[1].aid != track_two.alphabet.aid:
            s = "track {0} for sequence two has alphabet '{1}' but " \
                "the corresponding dimension in the score matrix " \
                "has alphabet '{2}'"
            raise DataError(s.format(0, track_two.alphabet.aid, score_matrix.alphabets[1].aid))
        for track in (track_one, track_two):
            if track.tid not in (PlainTrack.tid, ProfileTrack.tid):
                raise DataError("unknown track type id for this aligner: '{0}'".format(track.tid))
        sets.append((track_one, track_two, score_matrix))
    if len(gap_series) == 1:
        gaps = [gap_series[0], gap_series[0]]
    elif len(gap_series) == 2:
        gaps = list(gap_series)
    else:
        raise ComponentError("the fast aligner only supports linear and affine gap"
                             " penalties at the moment")
    return sets, gaps


def _profile_of(track):
    """component/align.py:163-172: one-hot f32 for plain tracks, .profile as f32 for profiles."""
    if track.tid == PlainTrack.tid:
        p = np.zeros((len(track), track.alphabet.size), dtype=np.float32)
        p[np.arange(len(track)), track.values] = 1.0
        return p
    return track.profile.astype(np.float32)


def _check_mode(mode):
    if mode not in MODES:
        raise ComponentError("unknown alignment mode: '{0}'".format(mode))


def _batchable(sets, gaps, zero_idxs, mode, engine):
    """One sequence-like track set, no mask, integer-exact scores: the inter-task kernel."""
    if len(sets) != 1 or zero_idxs:
        return None
    t1, t2, sm = sets[0]
    a, b = _seq_like(t1), _seq_like(t2)
    if a is None or b is None:
        return None
    S = sm.matrix.astype(np.float32)
    longest = max(len(a), len(b))
    if engine.k_for(longest) is None or not engine.integer_exact(S, gaps[0], gaps[1], longest):
        return None
    if mode == "local" and not (gaps[0] <= gaps[1] <= 0):
        # the reference takes argmax over the WHOLE o array, border cell o[0,0,1] = open - extend included
        # (component/align.py:371, :401-403); the batched local reduction covers interior cells only, which is the
        # same thing only while that border value is <= 0: anything else runs the general kernel
        return None
    return a, b, S


class GpuPairwiseAligner(Component):
    """GPU drop-in for praline.component.PairwiseAligner (component/align.py:37-251)."""
    tid = PAIRWISE_TID
    inputs = {'mode': Port(str),
              'sequence_one': Port(Sequence.tid),
              'sequence_two': Port(Sequence.tid),
              'track_id_sets_one': Port([[str]]),
              'track_id_sets_two': Port([[str]]),
              'zero_idxs': Port([(int, int)], optional=True),
              'score_matrices': Port([ScoreMatrix.tid])}
    outputs = {'alignment': Port(Alignment.tid), 'score': Port(float)}
    options = {'gap_series': [float], 'debug': int}
    defaults = {'gap_series': [-11.0, -1.0], 'debug': 0}

    def execute(self, mode, sequence_one, sequence_two, track_id_sets_one, track_id_sets_two, zero_idxs,
                score_matrices):
        gap_series = self.environment['gap_series']
        debug = self.environment['debug']
        log = None
        if debug > 0:       # the component's own log bundle, component/align.py:93-112
            log = LogBundle()
            log.message(ROOT_LOG_NAME, "Entering component '{0}'".format(self.tid))
            log.message(ROOT_LOG_NAME, "Alignment mode: '{0}'".format(mode))
            log.message(ROOT_LOG_NAME, "Gap scores: {0}".format(gap_series))
            msg = "Sequence one: '{0}', sequence two: '{1}'"
            log.message(ROOT_LOG_NAME, msg.format(sequence_one.name, sequence_two.name))
            log.message(ROOT_LOG_NAME, "Track id sets for sequence one:")
            for track_id_set in track_id_sets_one:
                log.message(ROOT_LOG_NAME, "\t{0}".format(", ".join(track_id_set)))
            log.message(ROOT_LOG_NAME, "Track id sets for sequence two:")
            for track_id_set in track_id_sets_two:
                log.message(ROOT_LOG_NAME, "\t{0}".format(", ".join(track_id_set)))
        sets, gaps = _prepare(sequence_one, sequence_two, track_id_sets_one, track_id_sets_two, score_matrices,
                              gap_series)
        _check_mode(mode)
        eng = get_engine()
        hit = _batchable(sets, gaps, zero_idxs, mode, eng) if (debug == 0 and mode != "local") else None
        if hit is not None:
            a, b, S = hit
            batch = eng.batch([a, b])
            scores, paths = eng.align_pairs(batch, [0], [1], S, gaps, mode=mode, want_paths=True, resident="two")
            score, path = float(scores[0]), paths[0]
            alignment = Alignment([sequence_one, sequence_two], _path_container(mode, path))
            # the message sequence of the reference: the nested RawPairwiseAligner task begins and completes under
            # this component's tag (component/align.py:225-237; bulk data stripped as Execution.run does)
            for msg in _nested_raw_messages(self.tag):
                yield msg
            yield CompleteMessage(outputs={'alignment': alignment, 'score': score})
            return
        # general path: score models on the device, then the raw aligner (component/align.py:200-237)
        p1s, p2s = [_profile_of(t1) for t1, _, _ in sets], [_profile_of(t2) for _, t2, _ in sets]
        mats = [sm.matrix.astype(np.float32) for _, _, sm in sets]
        L1, L2 = int(p1s[0].shape[0]), int(p2s[0].shape[0])
        if len(sets) == 1 and L1 * L2 >= (1 << 21):
            # one long profile pair: the raw aligner builds the scores beside its fill (pgpu_align_profile_long)
            m = (p1s[0], p2s[0], mats[0])
        else:
            m = eng.build_scores(p1s, p2s, mats)
        g1 = np.empty((L1, 2), dtype=np.float32)
        g2 = np.empty((L2, 2), dtype=np.float32)
        g1[:] = gaps
        g2[:] = gaps
        execution = Execution(self.manager, self.tag)
        task = execution.add_task(GpuRawPairwiseAligner)
        task.environment(self.environment)
        task.inputs(mode=mode, sequence_one=sequence_one, sequence_two=sequence_two,
                    match_score_model=DeviceMatchScoreModel(sequence_one, sequence_two, m),
                    gap_score_model_one=GapScoreModel(sequence_one, g1),
                    gap_score_model_two=GapScoreModel(sequence_two, g2), zero_idxs=zero_idxs)
        for msg in execution.run():
            yield msg
        outputs = execution.outputs[0]
        if log is not None:     # component/align.py:239-249
            log.message(ROOT_LOG_NAME, "Alignment score: {0}".format(outputs["score"]))
            log.message(ROOT_LOG_NAME, "Done!")
            archive_path = log.archive()
            log.delete()
            yield LogMessage(path_to_url(archive_path))
        yield CompleteMessage(outputs=outputs)


def _nested_raw_messages(parent_tag):
    """Begin + Complete of a RawPairwiseAligner task as a parent component sees them (core/manager.py:212-214,
    core/execution.py:176-186: tag = class name # uuid, outputs stripped on the way up)."""
    from uuid import uuid4
    tag = "{0}#{1}".format(RAW_TID.split(".")[-1], uuid4().hex)
    begin = BeginMessage(parent_tag)
    begin.tag = tag
    done = CompleteMessage(outputs=None)
    done.tag = tag
    return [begin, done]


class DeviceMatchScoreModel(MatchScoreModel):
    """A MatchScoreModel whose scores stay on the device between the score-matrix kernel and the
    fill (the reference materialises them on the host, component/align.py:205-219)."""
    tid = MatchScoreModel.tid

    def __init__(self, sequence_one, sequence_two, scores_dev):
        # scores_dev: the device matrix, or the operands (P1, P2, S) of ONE track set it is still to be built from
        self.operands = scores_dev if isinstance(scores_dev, tuple) else None
        shape = (self.operands[0].shape[0], self.operands[1].shape[0]) if self.operands is not None else scores_dev.shape
        if len(sequence_one) != shape[0]:
            s = "sequence length {0} does not correspond to array shape {1}"
            raise DataError(s.format(len(sequence_one), shape[0]))
        if len(sequence_two) != shape[1]:
            s = "sequence length {0} does not correspond to array shape {1}"
            raise DataError(s.format(len(sequence_two), shape[1]))
        self.sequence_one = sequence_one
        self.sequence_two = sequence_two
        self._scores_dev = None if self.operands is not None else scores_dev

    @property
    def scores_dev(self):
        if self._scores_dev is None:
            p1, p2, S = self.operands
            self._scores_dev = get_engine().build_scores([p1], [p2], [S])
        return self._scores_dev

    @property
    def scores(self):
        return self.scores_dev.cpu().numpy()


class GpuRawPairwiseAligner(Component):
    """GPU drop-in for praline.component.RawPairwiseAligner (component/align.py:254-447)."""
    tid = RAW_TID
    inputs = {'mode': Port(str),
              'sequence_one': Port(Sequence.tid),
              'sequence_two': Port(Sequence.tid),
              'match_score_model': Port(MatchScoreModel.tid),
              'gap_score_model_one': Port(GapScoreModel.tid),
              'gap_score_model_two': Port(GapScoreModel.tid),
              'zero_idxs': Port([(int, int)], optional=True)}
    outputs = {'alignment': Port(Alignment.tid), 'score': Port(float)}
    options = {'debug': int, 'accelerate': bool}
    defaults = {'debug': 0, 'accelerate': True}

    def execute(self, mode, sequence_one, sequence_two, match_score_model, gap_score_model_one,
                gap_score_model_two, zero_idxs):
        debug = self.environment['debug']
        if debug > 0:
            log = LogBundle()
            log.message(ROOT_LOG_NAME, "Entering component '{0}'".format(self.tid))
            log.message(ROOT_LOG_NAME, "Alignment mode: '{0}'".format(mode))
            msg = "Sequence one: '{0}', sequence two: '{1}'"
            log.message(ROOT_LOG_NAME, msg.format(sequence_one.name, sequence_two.name))
        _check_mode(mode)
        eng = get_engine()
        operands = getattr(match_score_model, "operands", None)
        if operands is not None and not zero_idxs and debug <= 1:
            # a long profile pair straight from GpuPairwiseAligner: K1 beside the fill, one library call
            r = eng.align_profile_pair(mode, operands[0], operands[1], operands[2], gap_score_model_one.scores,
                                       gap_score_model_two.scores)
        else:
            m = getattr(match_score_model, "scores_dev", None)
            if m is None:
                m = np.ascontiguousarray(match_score_model.scores, dtype=np.float32)
            r = eng.align_general(mode, m, gap_score_model_one.scores, gap_score_model_two.scores,
                                  zero_idxs=zero_idxs, want_matrices=debug > 1)
        if debug > 1:   # component/align.py:390-399
            log.message(ROOT_LOG_NAME, "Dumping DP & traceback matrices...")
            for k in range(3):
                np.savetxt(log.path("dp_{0}_matrix.csv".format(k)), r["o"][:, :, k], delimiter=",")
                np.savetxt(log.path("tb_{0}_matrix.csv".format(k)), r["t"][:, :, k], delimiter=",")
        alignment = Alignment([sequence_one, sequence_two], _path_container(mode, r["path"]))
        outputs = {'alignment': alignment, 'score': float(r["score"])}
        if debug > 0:
            log.message(ROOT_LOG_NAME, "Alignment score: {0}".format(r["score"]))
            log.message(ROOT_LOG_NAME, "Done!")
            archive_path = log.archive()
            log.delete()
            yield LogMessage(path_to_url(archive_path))
        yield CompleteMessage(outputs=outputs)


class GpuGuideTreeBuilder(Component):
    """GPU drop-in for praline.component.GuideTreeBuilder (component/tree.py:18-172): same type
    id, ports, options and defaults.  The reference queues one PairwiseAligner task per unordered
    pair (:97-131), fills the score matrix (:141-145), forms dist = -d + d.max() (:147) and runs
    the pure-Python clustering (util/cluster.py:27-57).  Here the N(N-1)/2 scores come from ONE
    all-vs-all launch (sequence tracks: the packed/f32 streaming kernel; profile tracks: the
    matrix-fed batch in the reference's evaluation order), the distance matrix stays on the
    device and the merge order comes from the clustering kernel (csrc/cluster.cu).  Anything the
    batched path does not cover (several track sets, semiglobal_auto, debug logging, a different
    aligner, over-long sequences) runs the reference's own component unchanged."""
    tid = GUIDE_TREE_TID

    inputs = {'sequences': Port([Sequence.tid]),
              'track_id_sets': Port([[str]]),
              'score_matrices': Port([ScoreMatrix.tid])}
    outputs = {'guide_tree': Port(SequenceTree.tid)}

    options = {'gap_series': [float], 'aligner': str,
               'aligner_env': Environment.tid,
               'linkage_method': str, 'squash_profiles': bool,
               'dist_mode': str, 'debug': int}
    defaults = {'gap_series': [-11.0, -1.0],
                'aligner': PAIRWISE_TID, 'aligner_env': Environment({}),
                'linkage_method': 'average', 'squash_profiles': False,
                'dist_mode': 'global', 'debug': 0}

    def _pair_scores(self, seqs, track_id_sets, score_matrices, mode):
        """Condensed scores of all pairs i < j (sequence_one = i, sequence_two = j, the order of
        tree.py:97-131) or None when the batched path does not apply."""
        env = self.environment
        if env['aligner'] != PAIRWISE_TID or self.manager.index.resolve(PAIRWISE_TID) is not GpuPairwiseAligner:
            return None
        if env['debug'] != 0 or len(track_id_sets) != 1 or len(seqs) < 2:
            return None
        sub_env = Environment(keys=env['aligner_env'].keys, component=GpuPairwiseAligner, parent=env)
        if sub_env['debug'] != 0:
            return None
        eng = get_engine()
        n = len(seqs)
        try:    # the checks of PairwiseAligner.execute on every sequence (against its neighbour)
            tracks = []
            for i in range(n):
                sets, gaps = _prepare(seqs[i], seqs[(i + 1) % n], track_id_sets, track_id_sets, score_matrices,
                                      sub_env['gap_series'])
                tracks.append(sets[0][0])
            sm = sets[0][2]
        except (ComponentError, DataError):
            return None     # the reference path raises it the reference's way
        S = sm.matrix.astype(np.float32)
        longest = max(len(t) for t in tracks)
        if min(len(t) for t in tracks) < 1 or eng.k_for(longest) is None:
            return None
        arrs = [_seq_like(t) for t in tracks]
        if all(a is not None for a in arrs) and eng.integer_exact(S, gaps[0], gaps[1], longest):
            batch = eng.batch(arrs)
            cond, _, _ = eng.allpairs_scores(batch, eng.dev(S), S.shape[0], gaps, mode=mode, S_host=S)
            return cond
        pi, pj = np.triu_indices(n, k=1)
        pb = eng.profile_batch([_profile_of(t) for t in tracks])
        import os
        fast = os.environ.get("PGPU_FAST_PROFILES", "") not in ("", "0")
        return eng.align_profile_pairs(pb, pi, pj, S, gaps, mode=mode, fast=fast)

    def execute(self, sequences, track_id_sets, score_matrices):
        linkage_method = self.environment['linkage_method']
        dist_mode = self.environment['dist_mode']
        if linkage_method not in {'single', 'complete', 'average'}:
            raise ComponentError("unknown linkage method '{0}'".format(linkage_method))
        if dist_mode not in {'semiglobal', 'global', 'semiglobal_auto'}:
            raise ComponentError("unknown alignment mode '{0}'".format(dist_mode))
        cond = None
        if dist_mode != 'semiglobal_auto':
            mode = "semiglobal_both" if dist_mode == "semiglobal" else "global"
            cond = self._pair_scores(sequences, track_id_sets, score_matrices, mode)
        if cond is None:
            from praline.component import GuideTreeBuilder as _ReferenceGuideTreeBuilder
            ref = _ReferenceGuideTreeBuilder(self.manager, self.environment, self.tag)
            for msg in ref.execute(sequences, track_id_sets, score_matrices):
                yield msg
            return
        eng = get_engine()
        n = len(sequences)
        yield ProgressMessage(1.0)
        dist = eng.tree_distance(cond, n)
        merges = eng.cluster_merge_order(dist, linkage_method)
        yield CompleteMessage({'guide_tree': SequenceTree(sequences, merges)})


def merge_profile_counts(counts_one, counts_two, path):
    """ProfileTrack.merge (container/sequence.py:205-239) without the per-column Python loop:
    column i of the merged profile is the sum of the count rows the path step i -> i+1 consumes
    (row path[i+1, 0] - 1 of profile one if index 0 advances, likewise for profile two).  The
    reference accumulates in f32 and ProfileTrack casts to int (sequence.py:184); the same here."""
    path = np.asarray(path)
    step = (path[1:] - path[:-1]) > 0
    merged = np.zeros((path.shape[0] - 1, counts_one.shape[1]), dtype=np.float32)
    r1 = np.flatnonzero(step[:, 0])
    merged[r1] += counts_one[path[r1 + 1, 0] - 1]
    r2 = np.flatnonzero(step[:, 1])
    merged[r2] += counts_two[path[r2 + 1, 1] - 1]
    return merged


def merge_alignment_paths(path_one, path_two, path):
    """The path of Alignment.merge (container/align.py:30-61): row i of the merged path is row
    path[i, 0] of alignment one next to row path[i, 1] of alignment two (-1 rows stay -1)."""
    path = np.asarray(path)
    one = np.where((path[:, 0] >= 0)[:, None], np.asarray(path_one)[np.maximum(path[:, 0], 0)], -1)
    two = np.where((path[:, 1] >= 0)[:, None], np.asarray(path_two)[np.maximum(path[:, 1], 0)], -1)
    return np.hstack([one, two]).astype(int)


class GpuTreeMultipleSequenceAligner(Component):
    """Drop-in for praline.component.TreeMultipleSequenceAligner (component/msa.py:22-247): same
    type id, ports, options and defaults.  The progressive merge is a dependency chain of
    profile x profile alignments (one per guide-tree node, msa.py:124-237), so it stays on one
    GPU; every alignment still goes through the manager as a PairwiseAligner task (the general
    K1 + K3 path), but the host glue between two launches -- ProfileTrack.merge and
    Alignment.merge, per-column Python loops in the reference -- is vectorised
    (merge_profile_counts, merge_alignment_paths).  Debug logging runs the reference component."""
    tid = TREE_MSA_TID

    inputs = {'sequences': Port([Sequence.tid]),
              'guide_tree': Port(SequenceTree.tid),
              'track_id_sets': Port([[str]]),
              'score_matrices': Port([ScoreMatrix.tid])}
    outputs = {'alignment': Port(Alignment.tid)}

    options = {'gap_series': [float], 'aligner': str,
               'aligner_env': Environment.tid, 'merge_mode': str,
               'debug': int, 'log_track_ids': [str]}
    defaults = {'gap_series': [-11.0, -1.0], 'aligner': PAIRWISE_TID,
                'aligner_env': Environment({}), 'merge_mode': 'semiglobal',
                'debug': 0, 'log_track_ids': [TRACK_ID_INPUT]}

    def _device_merge_plan(self, sequences, track_id_sets, score_matrices, merge_mode, track_ids):
        return _tree_device_merge_ok(self, sequences, track_id_sets, score_matrices, merge_mode, track_ids)

    def execute(self, sequences, guide_tree, track_id_sets, score_matrices):
        merge_mode = self.environment['merge_mode']
        if self.environment['debug'] > 0:
            from praline.component import TreeMultipleSequenceAligner as _Reference
            ref = _Reference(self.manager, self.environment, self.tag)
            for msg in ref.execute(sequences, guide_tree, track_id_sets, score_matrices):
                yield msg
            return
        if merge_mode not in {"global", "semiglobal", "semiglobal_auto"}:
            raise ComponentError("unknown merge mode '{0}'".format(merge_mode))
        track_ids = []
        for id_set in track_id_sets:
            for track_id in id_set:
                if track_id not in track_ids:
                    track_ids.append(track_id)
        # per cluster: the count profile of every track (msa.py:71-97) and the alignment path
        clusters = _cluster_tracks(sequences, track_ids)
        paths = {i: np.arange(len(seq) + 1).reshape(len(seq) + 1, 1) for i, seq in enumerate(sequences)}
        members = {i: [seq] for i, seq in enumerate(sequences)}
        index = self.manager.index
        total = len(guide_tree.merge_orders)
        device = self._device_merge_plan(sequences, track_id_sets, score_matrices, merge_mode, track_ids)
        if device is not None and total > 0:
            # Count tables stay on the device between merges and the independent merges of one guide-tree level run
            # together (Engine.merge_level): level(merge) = 1 + the deeper of its two clusters.  The alignment-path
            # bookkeeping (Alignment.merge) stays on the host -- it is the output.
            S, gaps, mode = device
            eng = get_engine()
            track_id = track_ids[0]
            A = S.shape[0]
            tables = {}
            for i in clusters:
                c = np.ascontiguousarray(clusters[i].get_track(track_id).counts, dtype=np.int32)
                tables[i] = (eng.dev(c), c.shape[0])
            depth = {i: 0 for i in clusters}
            levels = {}
            for (i, j) in guide_tree.merge_orders:
                d = max(depth[i], depth[j]) + 1
                depth[i] = d
                levels.setdefault(d, []).append((i, j))
            done = 0
            for d in sorted(levels):
                jobs = [(tables[i][0], tables[i][1], tables[j][0], tables[j][1]) for (i, j) in levels[d]]
                for (i, j), (merged, rows, path, _score) in zip(levels[d], eng.merge_level(jobs, S, gaps, mode)):
                    tables[i] = (merged, rows)
                    del tables[j]
                    paths[i] = merge_alignment_paths(paths[i], paths[j], path)
                    members[i] = members[i] + members[j]
                    del paths[j], members[j]
                    done += 1
                    yield ProgressMessage(progress=done / total)
            first = next(iter(paths))
            yield CompleteMessage(outputs={'alignment': Alignment(members[first], paths[first])})
            return
        for step, (i, j) in enumerate(guide_tree.merge_orders):
            one, two = clusters[i], clusters[j]
            if merge_mode == "semiglobal":
                mode = "semiglobal_both"
            elif merge_mode == "global":
                mode = "global"
            else:
                mode = auto_align_mode(one, two)
            execution = Execution(self.manager, self.tag)
            task = execution.add_task(index.resolve(self.environment['aligner']))
            task.environment(self.environment, self.environment['aligner_env'])
            task.inputs(mode=mode, sequence_one=one, sequence_two=two, track_id_sets_one=track_id_sets,
                        track_id_sets_two=track_id_sets, score_matrices=score_matrices)
            for msg in execution.run():
                yield msg
            path = np.array(execution.outputs[0]['alignment'].path)
            merged = []
            for track_id in track_ids:
                t1, t2 = one.get_track(track_id), two.get_track(track_id)
                if t2.tid != t1.tid:
                    raise DataError("can not merge with non-profile track {0}".format(t2.tid))
                merged.append((track_id, ProfileTrack(merge_profile_counts(t1.counts, t2.counts, path), t1.alphabet)))
                one.del_track(track_id)
            for track_id, track in merged:
                one.add_track(track_id, track)
            paths[i] = merge_alignment_paths(paths[i], paths[j], path)
            members[i] = members[i] + members[j]
            del clusters[j], paths[j], members[j]
            yield ProgressMessage(progress=(step + 1) / total)
        first = next(iter(paths))
        yield CompleteMessage(outputs={'alignment': Alignment(members[first], paths[first])})


def _tree_device_merge_ok(component, sequences, track_id_sets, score_matrices, merge_mode, track_ids):
    """The device-resident merge path of GpuTreeMultipleSequenceAligner serves one track set with one track, the
    GPU PairwiseAligner with constant gap penalties and merge_mode global / semiglobal; -> (S, gaps, mode) or None."""
    import os as _os
    env = component.environment
    if _os.environ.get("PGPU_NO_DEVICE_MERGE", "") != "" or merge_mode not in ("global", "semiglobal"):
        return None
    if len(track_id_sets) != 1 or len(track_id_sets[0]) != 1 or len(track_ids) != 1 or len(sequences) < 2:
        return None
    if env['aligner'] != PAIRWISE_TID or component.manager.index.resolve(PAIRWISE_TID) is not GpuPairwiseAligner:
        return None
    sub_env = Environment(keys=env['aligner_env'].keys, component=GpuPairwiseAligner, parent=env)
    if sub_env['debug'] != 0:
        return None
    try:    # the per-pair checks of PairwiseAligner.execute are per-sequence properties
        gaps = S = None
        for seq in sequences:
            sets, gaps = _prepare(seq, seq, track_id_sets, track_id_sets, score_matrices, sub_env['gap_series'])
            if len(sets) != 1:
                return None
            S = sets[0][2].matrix.astype(np.float32)
    except (ComponentError, DataError):
        return None
    track_id = track_id_sets[0][0]
    if min(len(seq.get_track(track_id)) for seq in sequences) < 1:
        return None
    return S, list(gaps), ("semiglobal_both" if merge_mode == "semiglobal" else "global")


def _cluster_tracks(sequences, track_ids):
    """Initial clusters of the MSA components (msa.py:71-97, :313-339): every track as a count
    profile (one-hot for plain tracks)."""
    clusters = {}
    for i, seq in enumerate(sequences):
        cluster = Sequence("Cluster #{0}".format(i), [])
        for track_id in track_ids:
            track = seq.get_track(track_id)
            if track.tid == PlainTrack.tid:
                counts = np.zeros((len(track), track.alphabet.size), dtype=np.int32)
                counts[np.arange(len(track)), np.asarray(track.values)] = 1
                cluster.add_track(track_id, ProfileTrack(counts, track.alphabet))
            elif track.tid == ProfileTrack.tid:
                cluster.add_track(track_id, ProfileTrack(np.array(track.counts, dtype=np.int32), track.alphabet))
        clusters[i] = cluster
    return clusters


class GpuAdHocMultipleSequenceAligner(Component):
    """Drop-in for praline.component.AdHocMultipleSequenceAligner (component/msa.py:249-560), the
    default MSA mode of the CLI (cmd.py:82-85): same type id, ports, options and defaults.

    The reference scores every pair of current clusters each round, reusing the scores that do
    not involve the cluster merged last (`_merge_indices`, msa.py:488-558), merges the first
    maximum of the score matrix in row-major order and repeats.  Here the score matrix lives in
    one array indexed by the original cluster ids (clusters keep their dict order, so list order
    is id order), the first round is ONE all-vs-all launch, every later round one batch of
    (merged cluster x every other cluster) profile alignments in the reference's orientation
    (sequence one = the cluster earlier in the list) against profiles that stay on the device
    (GrowingProfileBatch), and the merge itself is the vectorised glue of the tree aligner.  One
    track set, `dist_mode` global / semiglobal, the GPU PairwiseAligner and no debug logging;
    anything else runs the reference component."""
    tid = ADHOC_MSA_TID

    inputs = {'sequences': Port([Sequence.tid]),
              'track_id_sets': Port([[str]]),
              'score_matrices': Port([ScoreMatrix.tid])}
    outputs = {'alignment': Port(Alignment.tid)}

    options = {'gap_series': [float], 'aligner': str,
               'aligner_env': Environment.tid, 'merge_mode': str,
               'dist_mode': str, 'debug': int, 'log_track_ids': [str]}
    defaults = {'gap_series': [-11.0, -1.0], 'aligner': PAIRWISE_TID,
                'aligner_env': Environment({}), 'merge_mode': 'semiglobal',
                'dist_mode': 'global', 'debug': 0,
                'log_track_ids': [TRACK_ID_INPUT]}

    def _reference(self, sequences, track_id_sets, score_matrices):
        from praline.component import AdHocMultipleSequenceAligner as _Reference
        ref = _Reference(self.manager, self.environment, self.tag)
        for msg in ref.execute(sequences, track_id_sets, score_matrices):
            yield msg

    def execute(self, sequences, track_id_sets, score_matrices):
        env = self.environment
        merge_mode, dist_mode = env['merge_mode'], env['dist_mode']
        if merge_mode not in {'global', 'semiglobal', 'semiglobal_auto'}:
            raise ComponentError("unknown merge mode '{0}'".format(merge_mode))
        if dist_mode not in {'global', 'semiglobal', 'semiglobal_auto'}:
            raise ComponentError("unknown distance mode '{0}'".format(dist_mode))
        eng = get_engine()
        ok = (env['debug'] == 0 and dist_mode != 'semiglobal_auto' and len(track_id_sets) == 1 and
              len(track_id_sets[0]) == 1 and len(sequences) >= 2 and env['aligner'] == PAIRWISE_TID and
              self.manager.index.resolve(PAIRWISE_TID) is GpuPairwiseAligner)
        sub_env = gaps = S = None
        if ok:
            sub_env = Environment(keys=env['aligner_env'].keys, component=GpuPairwiseAligner, parent=env)
            ok = sub_env['debug'] == 0
        track_id = track_id_sets[0][0] if ok else None
        if ok:
            try:    # the per-pair checks of PairwiseAligner.execute are per-sequence properties
                for seq in sequences:
                    sets, gaps = _prepare(seq, seq, track_id_sets, track_id_sets, score_matrices, sub_env['gap_series'])
                S = sets[0][2].matrix.astype(np.float32)
            except (ComponentError, DataError):
                ok = False
        if ok:
            lens = [len(seq.get_track(track_id)) for seq in sequences]
            ok = min(lens) >= 1 and eng.k_for(max(lens)) is not None
        if not ok:
            for msg in self._reference(sequences, track_id_sets, score_matrices):
                yield msg
            return

        dmode = "semiglobal_both" if dist_mode == "semiglobal" else "global"
        n = len(sequences)
        clusters = _cluster_tracks(sequences, [track_id])
        paths = {i: np.arange(len(seq) + 1).reshape(len(seq) + 1, 1) for i, seq in enumerate(sequences)}
        members = {i: [seq] for i, seq in enumerate(sequences)}
        import os
        fast = os.environ.get("PGPU_FAST_PROFILES", "") not in ("", "0")
        # score matrix over the original ids, upper triangle; dead or unset entries are -inf, so its
        # first row-major maximum is the reference's (msa.py:549: first maximum of the symmetric s)
        sc = np.full((n, n), -np.inf, dtype=np.float32)
        tracks = [clusters[i].get_track(track_id) for i in range(n)]
        arrs = [_seq_like(t) for t in tracks]
        iu = np.triu_indices(n, k=1)
        pb = eng.profile_batch([_profile_of(t) for t in tracks], cap_rows=3 * sum(lens))
        pidx = {i: i for i in range(n)}             # cluster id -> its current profile in the batch
        if all(a is not None for a in arrs) and eng.integer_exact(S, gaps[0], gaps[1], max(lens)) and \
                eng.k_for(max(lens)) is not None:
            batch = eng.batch(arrs)
            cond, _, _ = eng.allpairs_scores(batch, eng.dev(S), S.shape[0], gaps, mode=dmode, S_host=S)
            sc[iu] = cond.cpu().numpy()
        else:
            sc[iu] = eng.align_profile_pairs(pb, iu[0], iu[1], S, gaps, mode=dmode, fast=fast)
        alive = list(range(n))
        index = self.manager.index
        total = n - 1
        for step in range(total):
            i, j = np.unravel_index(sc.argmax(), sc.shape)
            i, j = int(i), int(j)
            one, two = clusters[i], clusters[j]
            if merge_mode == "semiglobal":
                mode = "semiglobal_both"
            elif merge_mode == "global":
                mode = "global"
            else:
                mode = auto_align_mode(one, two)
            execution = Execution(self.manager, self.tag)
            task = execution.add_task(index.resolve(env['aligner']))
            task.environment(env, env['aligner_env'])
            task.inputs(mode=mode, sequence_one=one, sequence_two=two, track_id_sets_one=track_id_sets,
                        track_id_sets_two=track_id_sets, score_matrices=score_matrices)
            for msg in execution.run():
                yield msg
            path = np.array(execution.outputs[0]['alignment'].path)
            t1, t2 = one.get_track(track_id), two.get_track(track_id)
            merged = ProfileTrack(merge_profile_counts(t1.counts, t2.counts, path), t1.alphabet)
            one.del_track(track_id)
            one.add_track(track_id, merged)
            paths[i] = merge_alignment_paths(paths[i], paths[j], path)
            members[i] = members[i] + members[j]
            del clusters[j], paths[j], members[j]
            alive.remove(j)
            sc[j, :] = -np.inf
            sc[:, j] = -np.inf
            yield ProgressMessage(progress=(step + 1) / total)
            if len(alive) < 2:
                break
            # the merged cluster against every other one, sequence one = the earlier cluster
            pidx[i] = pb.append(_profile_of(merged))
            before = [k for k in alive if k < i]
            after = [k for k in alive if k > i]
            if eng.k_for(len(merged)) is None:
                # a cluster wider than the inter-task kernels: its pairs go through the general path
                prof = {k: _profile_of(clusters[k].get_track(track_id)) for k in alive}
                for a_, b_ in [(i, k) for k in after] + [(k, i) for k in before]:
                    m = eng.build_scores([prof[a_]], [prof[b_]], [S])
                    g1 = np.empty((prof[a_].shape[0], 2), np.float32)
                    g2 = np.empty((prof[b_].shape[0], 2), np.float32)
                    g1[:] = gaps
                    g2[:] = gaps
                    sc[a_, b_] = eng.align_general(dmode, m, g1, g2, want_path=False)["score"]
                continue
            if after:       # (i, k): the merged cluster is sequence one and the resident of all its pairs
                got = eng.align_profile_pairs(pb, [pidx[i]] * len(after), [pidx[k] for k in after], S, gaps,
                                              mode=dmode, resident="one", fast=fast)
                sc[i, after] = got
            if before:      # (k, i): the merged cluster is sequence two, still the shared resident
                got = eng.align_profile_pairs(pb, [pidx[k] for k in before], [pidx[i]] * len(before), S, gaps,
                                              mode=dmode, resident="two", fast=fast)
                sc[before, i] = got
        first = next(iter(paths))
        yield CompleteMessage(outputs={'alignment': Alignment(members[first], paths[first])})


def register(index, tree=True):
    """Replace the CPU aligners and (tree=True) the guide-tree builder of a TypeIndex
    (manager.py:49-57) by the GPU ones."""
    index.register(GpuPairwiseAligner)
    index.register(GpuRawPairwiseAligner)
    if tree:
        index.register(GpuGuideTreeBuilder)
        index.register(GpuTreeMultipleSequenceAligner)
        index.register(GpuAdHocMultipleSequenceAligner)
    return index


class _Group(object):
    """Pairs of one (mode, matrix, gaps) class collected from a request list."""

    def __init__(self, mode, S, gaps):
        self.mode, self.S, self.gaps = mode, S, gaps
        self.seq_index = {}
        self.seqs = []
        self.pi, self.pj = [], []
        self._scores = None
        self.paths = None
        self._batch = None

    def add_seq(self, track, arr):
        k = id(track)
        if k not in self.seq_index:
            self.seq_index[k] = len(self.seqs)
            self.seqs.append(arr)
        return self.seq_index[k]

    def add_pair(self, t1, a, t2, b):
        self.pi.append(self.add_seq(t1, a))
        self.pj.append(self.add_seq(t2, b))
        return len(self.pi) - 1

    def add_pairs(self, one_idx, two_idxs):
        """Bulk form: sequence one against many sequences two; returns the first pair number."""
        k0 = len(self.pi)
        self.pi.extend([one_idx] * len(two_idxs))
        self.pj.extend(two_idxs)
        return k0

    @property
    def batch(self):
        if self._batch is None:
            self._batch = get_engine().batch(self.seqs)
        return self._batch

    @property
    def scores(self):
        """Score per pair, computed by one score-only launch on first use."""
        if self._scores is None:
            self._scores, _ = get_engine().align_pairs(self.batch, self.pi, self.pj, self.S, self.gaps, mode=self.mode)
        return self._scores

    def run_scores(self, eng):
        self.scores

    def path(self, k):
        if self.paths is None:   # first access traces the whole group at once
            eng = get_engine()
            if self.mode == "local":
                self.paths = {}
            else:
                _, self.paths = eng.align_pairs(self.batch, self.pi, self.pj, self.S, self.gaps, mode=self.mode,
                                                want_paths=True)
        if self.mode == "local":
            if k not in self.paths:
                r = get_engine().align_seq_pair_general(self.batch, self.pi[k], self.pj[k], self.S, self.gaps,
                                                        self.mode)
                self.paths[k] = r["path"]
            return _path_container(self.mode, self.paths[k])
        return _path_container(self.mode, self.paths[k])


class _ProfileGroup(object):
    """Profile x profile pairs of one (mode, matrix, gaps) class: scores from one matrix-fed
    batch launch, paths (rarely asked for) from the exact general path, pair by pair."""

    def __init__(self, mode, S, gaps):
        self.mode, self.S, self.gaps = mode, S, gaps
        self.index = {}
        self.profiles = []
        self.pi, self.pj = [], []
        self.scores = None
        self.paths = {}

    def add_profile(self, track):
        k = id(track)
        if k not in self.index:
            self.index[k] = len(self.profiles)
            self.profiles.append(_profile_of(track))
        return self.index[k]

    def add_pair(self, t1, a, t2, b):
        self.pi.append(self.add_profile(t1))
        self.pj.append(self.add_profile(t2))
        return len(self.pi) - 1

    def run_scores(self, eng):
        # PGPU_FAST_PROFILES=1 opts into the tolerance-mode score rows (<= 1e-5 relative, A FMAs per
        # cell); the default keeps the reference's evaluation order, bit for bit
        import os
        fast = os.environ.get("PGPU_FAST_PROFILES", "") not in ("", "0")
        pb = eng.profile_batch(self.profiles)
        self.scores = eng.align_profile_pairs(pb, self.pi, self.pj, self.S, self.gaps, mode=self.mode, fast=fast)

    def path(self, k):
        if k not in self.paths:
            eng = get_engine()
            p1, p2 = self.profiles[self.pi[k]], self.profiles[self.pj[k]]
            m = eng.build_scores([p1], [p2], [self.S])
            g1 = np.empty((p1.shape[0], 2), np.float32)
            g2 = np.empty((p2.shape[0], 2), np.float32)
            g1[:] = self.gaps
            g2[:] = self.gaps
            self.paths[k] = eng.align_general(self.mode, m, g1, g2)["path"]
        return _path_container(self.mode, self.paths[k])


class GpuBatchManager(Manager):
    """Manager whose execute_many aligns all PairwiseAligner requests of an Execution in one
    batched launch per (mode, matrix, gaps) group (manager.py:154-170)."""

    def __init__(self, index, register_gpu=True, gpu_tree=True):
        super(GpuBatchManager, self).__init__(register(index, tree=gpu_tree) if register_gpu else index)
        self.batched_requests = 0

    def execute_many(self, requests, parent_tag):
        if not self.open:
            from praline.core import PralineError
            raise PralineError("manager has been closed")
        requests = list(requests)
        plain, ms, lms = [], [], []
        for n, (tid, inputs, tag, env) in enumerate(requests):
            if tid == PAIRWISE_TID and self.index.resolve(tid) is GpuPairwiseAligner:
                plain.append(n)
            elif tid == GLOBAL_MS_TID and len(requests) > 0:
                ms.append(n)
            elif tid == LOCAL_MS_TID:
                lms.append(n)
        handled = set()
        if plain:
            for msg in self._batched_pairwise(requests, plain, parent_tag, handled):
                yield msg
        if ms:
            for msg in self._batched_master_slave(requests, ms, parent_tag, handled):
                yield msg
        if lms and self._bulk_master_slave(requests, lms, parent_tag, handled, ms_tid=LOCAL_MS_TID):
            for msg in self._bulk_messages(parent_tag, handled):
                yield msg
        for n, (tid, inputs, tag, env) in enumerate(requests):
            if n in handled:
                continue
            if tid == PROFILE_BUILDER_TID:
                done = self._device_profile(inputs, env)
                if done is not None:
                    begin = BeginMessage(parent_tag)
                    begin.tag = tag
                    yield begin
                    done.tag = tag
                    yield done
                    continue
            for msg in self._invoke(tid, inputs, tag, env, parent_tag=parent_tag):
                yield msg

    def _device_profile(self, inputs, env):
        """ProfileBuilder (profile.py:41-74) on a lazily built master-slave alignment: the count
        table comes from the device, the alignment itself is never materialised."""
        alignment = inputs.get('alignment')
        if not isinstance(alignment, LazyMasterSlaveAlignment):
            return None
        comp = self.index.resolve(PROFILE_BUILDER_TID)
        if Environment(keys=env.keys, component=comp)['debug'] != 0:
            return None
        freqs = alignment.gpu_counts(inputs.get('track_id'))
        if freqs is None:
            return None
        track = alignment._master.get_track(inputs['track_id'])
        return CompleteMessage(outputs={'profile_track': ProfileTrack(freqs, track.alphabet)})

    # -- PairwiseAligner requests ----------------------------------------------------------------
    def _collect(self, tid, inputs, env, groups, eng):
        """Validate one PairwiseAligner request like Manager._invoke + PairwiseAligner.execute do and
        file it into a group; returns (group, k) or None when it needs the general path."""
        comp = GpuPairwiseAligner
        env = Environment(keys=env.keys, component=comp) if 'gap_series' not in env.keys else env
        if env['debug'] != 0 or inputs.get('zero_idxs'):
            return None
        mode = inputs['mode']
        sets, gaps = _prepare(inputs['sequence_one'], inputs['sequence_two'], inputs['track_id_sets_one'],
                              inputs['track_id_sets_two'], inputs['score_matrices'], env['gap_series'])
        _check_mode(mode)
        hit = _batchable(sets, gaps, None, mode, eng)
        if hit is None:
            # profile x profile, one track set: matrix-fed batch (scores), exact general path (paths)
            if len(sets) != 1:
                return None
            if mode == "local" and not (gaps[0] <= gaps[1] <= 0):
                return None     # a positive border cell open - extend takes part in the reference's argmax: general kernel
            t1, t2, sm = sets[0]
            if eng.k_for(max(len(t1), len(t2))) is None:
                return None
            key = ("prof", mode, id(sm), float(gaps[0]), float(gaps[1]))
            if key not in groups:
                groups[key] = _ProfileGroup(mode, sm.matrix.astype(np.float32), gaps)
            g = groups[key]
            return g, g.add_pair(t1, None, t2, None)
        a, b, S = hit
        key = (mode, id(sets[0][2]), float(gaps[0]), float(gaps[1]))
        if key not in groups:
            groups[key] = _Group(mode, S, gaps)
        g = groups[key]
        return g, g.add_pair(sets[0][0], a, sets[0][1], b)

    def _batched_pairwise(self, requests, idxs, parent_tag, handled):
        eng = get_engine()
        groups, where = {}, {}
        for n in idxs:
            tid, inputs, tag, env = requests[n]
            hit = self._collect(tid, inputs, env, groups, eng)
            if hit is not None:
                where[n] = hit
        if len(where) < 2:
            return
        for g in groups.values():
            g.run_scores(eng)
        for n in idxs:
            if n not in where:
                continue
            tid, inputs, tag, env = requests[n]
            g, k = where[n]
            begin = BeginMessage(parent_tag)
            begin.tag = tag
            yield begin
            for msg in _nested_raw_messages(tag):       # the reference's nested RawPairwiseAligner task
                yield msg
            alignment = LazyAlignment([inputs['sequence_one'], inputs['sequence_two']], g.path, k)
            done = CompleteMessage(outputs={'alignment': alignment, 'score': float(g.scores[k])})
            done.tag = tag
            yield done
            handled.add(n)
            self.batched_requests += 1

    # -- GlobalMasterSlaveAligner requests (preprofile.py:67-156) --------------------------------
    def _bulk_messages(self, parent_tag, handled):
        for n, tag, alignment, npairs in self._bulk_out:
            begin = BeginMessage(parent_tag)
            begin.tag = tag
            yield begin
            prog = ProgressMessage(1.0)
            prog.tag = tag
            yield prog
            done = CompleteMessage({'alignment': alignment})
            done.tag = tag
            yield done
            handled.add(n)
            self.batched_requests += npairs
        self._bulk_out = []

    def _bulk_master_slave(self, requests, idxs, parent_tag, handled, ms_tid=GLOBAL_MS_TID):
        """The common shape of the preprofile stage (workflow.py:139-161): N requests over ONE set of
        plain-track sequences, one track set, one matrix, one aligner environment.  Every sequence
        is validated once (the per-pair checks of PairwiseAligner.execute, component/align.py:114-189,
        are per-sequence properties) and the N(N-1) ordered pairs are filed with list operations
        only -- no per-pair Python.  Returns False when the requests do not have that shape.

        ms_tid = LOCAL_MS_TID: the same for LocalMasterSlaveAligner (preprofile.py:160-267): local
        alignments with Waterman-Eggert iterations, whose bounding-box masks stay on the device
        (Engine.local_pairs); every (pair, iteration) above the threshold adds into the master's table."""
        local = ms_tid == LOCAL_MS_TID
        eng = get_engine()
        if self.index.resolve(PAIRWISE_TID) is not GpuPairwiseAligner:
            return False
        first = requests[idxs[0]][1]
        track_ids, mats = first['track_id_sets'], first['score_matrices']
        if len(track_ids) != 1 or len(track_ids[0]) != 1 or len(mats) < 1:
            return False
        tid0 = track_ids[0][0]
        comp = self.index.resolve(ms_tid)
        envs = []
        key0 = None
        for n in idxs:
            _, inputs, _, env0 = requests[n]
            if inputs['track_id_sets'] != track_ids or len(inputs['score_matrices']) != len(mats) or \
                    any(a is not b for a, b in zip(inputs['score_matrices'], mats)):
                return False
            env = Environment(keys=env0.keys, component=comp)
            if env['aligner'] != PAIRWISE_TID:
                return False
            sub_env = Environment(keys=env['aligner_env'].keys, component=GpuPairwiseAligner, parent=env)
            key = (tuple(sub_env['gap_series']), sub_env['debug'], env['score_threshold'],
                   env['waterman_eggert_iterations'] if local else None)
            if key0 is None:
                key0 = key
            if key != key0 or sub_env['debug'] != 0:
                return False
            envs.append(env)
        gap_series, _, threshold, iterations = key0
        if local and not 1 <= iterations <= eng.NBOX + 1:
            return False
        # one validation per distinct sequence
        seq_idx, tracks = {}, []
        gaps = None

        def visit(seq):
            if id(seq) in seq_idx:
                return True
            sets, g_ = _prepare(seq, seq, track_ids, track_ids, mats, list(gap_series))
            if sets[0][0].tid != PlainTrack.tid:
                return False
            seq_idx[id(seq)] = len(tracks)
            tracks.append(sets[0][0])
            return True

        # the workflow's own shape (workflow.py:66-73): request i's master is sequence i, its slaves are
        # all the others in order -- checked with C-speed list comparisons, no per-pair Python at all
        everyone = [requests[n][1]['master_sequence'] for n in idxs]
        try:
            _, gaps = _prepare(everyone[0], everyone[0], track_ids, track_ids, mats, list(gap_series))
            if not all(visit(seq) for seq in everyone):
                return False
            complete = len(tracks) == len(everyone) > 1 and \
                all(requests[n][1]['slave_sequences'] == everyone[:i] + everyone[i + 1:] for i, n in enumerate(idxs))
            if not complete:
                for n in idxs:
                    if not all(visit(seq) for seq in requests[n][1]['slave_sequences']):
                        return False
        except (ComponentError, DataError):
            return False        # the per-request path reports it the reference's way
        arrs = [np.asarray(t.values) for t in tracks]
        S = mats[0].matrix.astype(np.float32)
        longest = max(len(a) for a in arrs)
        if min(len(a) for a in arrs) < 1 or eng.k_for(longest) is None or \
                not eng.integer_exact(S, gaps[0], gaps[1], longest):
            return False
        if local and not eng.local_batchable(S, gaps, [len(a) for a in arrs]):
            return False
        g = _Group("local" if local else "global", S, gaps)
        for t, a in zip(tracks, arrs):
            g.add_seq(t, a)
        self._bulk_out = []
        if complete:
            pre = _StagePreprofile(g, threshold, iterations if local else None)
            for i, n in enumerate(idxs):
                _, inputs, tag, _ = requests[n]
                alignment = LazyMasterSlaveAlignment(everyone[i], inputs['slave_sequences'], None, threshold, tid0, pre, i)
                self._bulk_out.append((n, tag, alignment, (len(everyone) - 1) * (iterations if local else 1)))
            return True
        pre = _PreprofileBatch(g, threshold, iterations if local else None)
        for n, env in zip(idxs, envs):
            _, inputs, tag, _ = requests[n]
            master, slaves = inputs['master_sequence'], inputs['slave_sequences']
            midx = seq_idx[id(master)]
            sidx = [seq_idx[id(s_)] for s_ in slaves]
            k0 = g.add_pairs(midx, sidx)
            pre.add_many(midx, sidx)
            alignment = LazyMasterSlaveAlignment(master, slaves, _RangePicks(g, k0, len(sidx)), threshold, tid0, pre, midx)
            self._bulk_out.append((n, tag, alignment, len(sidx) * (iterations if local else 1)))
        return True

    def _batched_master_slave(self, requests, idxs, parent_tag, handled):
        if len(idxs) > 1 and self._bulk_master_slave(requests, idxs, parent_tag, handled):
            for msg in self._bulk_messages(parent_tag, handled):
                yield msg
            return
        eng = get_engine()
        groups, plans = {}, {}
        for n in idxs:
            tid, inputs, tag, env = requests[n]
            comp = self.index.resolve(tid)
            env = Environment(keys=env.keys, component=comp)
            if env['aligner'] != PAIRWISE_TID or self.index.resolve(PAIRWISE_TID) is not GpuPairwiseAligner:
                continue
            sub_env = Environment(keys=env['aligner_env'].keys, component=GpuPairwiseAligner, parent=env)
            master = inputs['master_sequence']
            slaves = inputs['slave_sequences']
            items = []
            ok = True
            for slave in slaves:
                sub_inputs = dict(mode="global", sequence_one=master, sequence_two=slave,
                                  track_id_sets_one=inputs['track_id_sets'], track_id_sets_two=inputs['track_id_sets'],
                                  score_matrices=inputs['score_matrices'])
                hit = self._collect(PAIRWISE_TID, sub_inputs, sub_env, groups, eng)
                if hit is None:
                    ok = False
                    break
                items.append(hit)
            if ok:
                plans[n] = (env, items)
        if not plans:
            return
        for g in groups.values():
            g.run_scores(eng)
        pre = {}
        for n in idxs:
            if n not in plans:
                continue
            tid, inputs, tag, env0 = requests[n]
            env, items = plans[n]
            begin = BeginMessage(parent_tag)
            begin.tag = tag
            yield begin
            master, slaves = inputs['master_sequence'], inputs['slave_sequences']
            threshold = env['score_threshold']
            track_ids = inputs['track_id_sets']
            plain = len(track_ids) == 1 and all(s_.get_track(track_ids[0][0]).tid == PlainTrack.tid
                                                for s_ in [master] + list(slaves))
            if plain and items and isinstance(items[0][0], _Group):
                g = items[0][0]
                key = (id(g), threshold)
                if key not in pre:
                    pre[key] = _PreprofileBatch(g, threshold)
                midx = g.pi[items[0][1]]
                for (gg, k) in items:
                    pre[key].add(gg.pi[k], gg.pj[k])
                alignment = LazyMasterSlaveAlignment(master, slaves, items, threshold, track_ids[0][0], pre[key], midx)
            else:   # not plain tracks: the reference's host-side merge, paths from the traced batch
                alignment = LazyMasterSlaveAlignment(master, slaves, items, threshold, None, None, None)._build()
            prog = ProgressMessage(1.0)
            prog.tag = tag
            yield prog
            done = CompleteMessage({'alignment': alignment})
            done.tag = tag
            yield done
            handled.add(n)
            self.batched_requests += len(items)
