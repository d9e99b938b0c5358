#!/usr/bin/env python
"""Golden vectors for the local master-slave stage, made by RUNNING THE REFERENCE's components
(Manager -> LocalMasterSlaveAligner -> PairwiseAligner(mode="local", zero_idxs=...) and
ProfileBuilder; praline/component/preprofile.py:160-267, profile.py:41-74) on small seeded families.
Needs the reference in baseline/_ref (build container only).

Per case: the sequences, the options, and for every master
  * every PairwiseAligner call the component made (slave, iteration, score, path) -- recorded by a
    manager that looks at the CompleteMessages of the sub-executions,
  * the master-slave alignment path and the ProfileBuilder count table.

    python tests/golden/make_local_ms_golden.py   ->  tests/golden/local_ms.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_praline as R  # noqa: E402
import praline  # noqa: E402
import praline.component as pc  # noqa: E402
from praline.core import Manager, Execution, Environment  # noqa: E402
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT  # noqa: E402
from praline_b200 import synth  # noqa: E402


class RecordingManager(Manager):
    """Stock Manager that keeps (score, path) of every PairwiseAligner it runs."""

    def __init__(self, index):
        Manager.__init__(self, index)
        self.calls = []

    def _invoke(self, tid, inputs, tag, env, parent_tag=None):
        for msg in Manager._invoke(self, tid, inputs, tag, env, parent_tag=parent_tag):
            if tid == pc.PairwiseAligner.tid and getattr(msg, "outputs", None) is not None and 'score' in msg.outputs:
                self.calls.append((inputs['sequence_two'].name, len(inputs.get('zero_idxs') or []),
                                   float(msg.outputs['score']),
                                   [[int(a), int(b)] for a, b in msg.outputs['alignment'].path]))
            yield msg


def seq(name, idx):
    return Sequence(name, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=np.asarray(idx)))])


with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)

cases = []
rng = np.random.default_rng(77)
specs = [
    dict(seed=41, n=5, length=40, iterations=2, threshold=None, gaps=[-11.0, -1.0], extra=0),
    dict(seed=42, n=6, length=55, iterations=3, threshold=40.0, gaps=[-11.0, -1.0], extra=1),
    dict(seed=43, n=4, length=33, iterations=4, threshold=None, gaps=[-4.0], extra=1),
    dict(seed=44, n=5, length=70, iterations=1, threshold=None, gaps=[-8.0, -2.0], extra=0),
    dict(seed=45, n=4, length=48, iterations=2, threshold=1000.0, gaps=[-11.0, -1.0], extra=0),
]
for sp in specs:
    fam = [np.asarray(s) for s in synth.family(sp["seed"], sp["n"], sp["length"])]
    for e in range(sp["extra"]):     # an unrelated sequence: short, low-scoring local alignments
        fam.append(rng.integers(0, 20, 17 + 9 * e).astype(np.int32))
    seqs = [seq("s%d" % i, s) for i, s in enumerate(fam)]
    case = dict(sp, seqs=[s.tolist() for s in fam], masters=[])
    for i, master in enumerate(seqs):
        mgr = RecordingManager(R.reference_index())
        keys = {'gap_series': sp["gaps"], 'score_threshold': sp["threshold"], 'aligner': pc.PairwiseAligner.tid,
                'waterman_eggert_iterations': sp["iterations"]}
        slaves = [s for j, s in enumerate(seqs) if j != i]
        out, _ = R.run_task(mgr, pc.LocalMasterSlaveAligner, keys, master_sequence=master, slave_sequences=slaves,
                            track_id_sets=[[TRACK_ID_INPUT]], score_matrices=[sm])
        aln = out['alignment']
        prof, _ = R.run_task(Manager(R.reference_index()), pc.ProfileBuilder, {}, alignment=aln, track_id=TRACK_ID_INPUT)
        case["masters"].append(dict(
            calls=[dict(slave=c[0], n_zero=c[1], score=c[2], path=c[3]) for c in mgr.calls],
            items=[it.name for it in aln.items], path=np.asarray(aln.path).tolist(),
            counts=np.asarray(prof['profile_track'].counts).tolist()))
    cases.append(case)
with open(os.path.join(HERE, "local_ms.json"), "w") as f:
    json.dump(cases, f)
print("wrote %d cases, %d alignments" % (len(cases), sum(len(m["calls"]) for c in cases for m in c["masters"])))
