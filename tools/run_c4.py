"""BASELINE config 4: the reference's MSA workflow (global preprofiles, guide tree,
progressive merge with merge_mode='semiglobal' -> semiglobal_both profile x profile
alignments on f32 profiles) on GpuBatchManager at full size, and the stock single-process
Manager on a subsample for the CPU figure (labelled; the full CPU run is hours).

    python tools/run_c4.py [n_seqs=2000] [length=400] [cpu_subsample=60]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ref_praline as R
from praline_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sub = int(sys.argv[3]) if len(sys.argv) > 3 else 60

import praline
from praline.core import Manager
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT
from praline_b200 import plugin, get_engine

with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
fam = synth.family(4, n, L)
mk = lambda k: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=s))])
                for i, s in enumerate(fam[:k])]
extra = {'merge_mode': 'semiglobal'}
res = {"stage": "C4 progressive merge workflow", "n_seqs": n, "length": L, "merge_mode": "semiglobal",
       "fast_profiles": os.environ.get("PGPU_FAST_PROFILES", "0"), "cores": os.cpu_count()}
mgr = plugin.GpuBatchManager(R.reference_index())
R.workflow_fasta(mgr, mk(4), sm, "global", "tree", extra=extra)          # warm-up
eng = get_engine()
l0 = eng.launches
t0 = time.perf_counter()
got = R.workflow_fasta(mgr, mk(n), sm, "global", "tree", extra=extra)
res["gpu_s"] = time.perf_counter() - t0
res["gpu_launches"] = eng.launches - l0
import hashlib
res["alignment_columns"] = len(got.split("\n")[1]) if got else 0
res["fasta_md5"] = hashlib.md5(got.encode()).hexdigest() if got else None
if sub > 1:
    mgr2 = plugin.GpuBatchManager(R.reference_index())
    t0 = time.perf_counter()
    g2 = R.workflow_fasta(mgr2, mk(sub), sm, "global", "tree", extra=extra)
    res["gpu_s_subsample"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = R.workflow_fasta(Manager(R.reference_index()), mk(sub), sm, "global", "tree", extra=extra)
    res["cpu_1core_s_subsample"] = time.perf_counter() - t0
    res["subsample"] = sub
    res["identical_subsample"] = bool(g2 == want)
print(json.dumps(res))
