"""End-to-end MSA wall time (BASELINE metric, second half): the reference's
PralineMultipleSequenceAlignmentWorkflow (what `praline in.fa out.aln --preprofile-global
--msa-tree` runs, praline/cmd.py:56-119) on the stock CPU Manager vs the same workflow on
GpuBatchManager.  Outputs must be byte-identical.  Prints one JSON line.

    python tools/msa_e2e.py [n_seqs] [length] [preprofile] [msa]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ref_praline as R
from praline_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pre = sys.argv[3] if len(sys.argv) > 3 else "global"
msa = sys.argv[4] if len(sys.argv) > 4 else "tree"
skip_cpu = len(sys.argv) > 5 and sys.argv[5] == "nocpu"

import praline
from praline.core import Manager
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT
from praline_b200 import plugin

with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
fam = synth.family(1, n, L)
mk = lambda: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=s))]) for i, s in enumerate(fam)]

res = {"n_seqs": n, "length": L, "preprofile": pre, "msa": msa, "cores": os.cpu_count()}
mgr = plugin.GpuBatchManager(R.reference_index())
R.workflow_fasta(mgr, mk()[:4], sm, pre, msa)          # warm-up: CUDA context, kernels
t0 = time.perf_counter()
got = R.workflow_fasta(mgr, mk(), sm, pre, msa)
res["gpu_first_s"] = time.perf_counter() - t0      # first full-size run of the process: device allocator and pinned staging warm up
t0 = time.perf_counter()
got2 = R.workflow_fasta(mgr, mk(), sm, pre, msa)
res["gpu_s"] = time.perf_counter() - t0            # steady state (same process, second run)
assert got2 == got
res["gpu_batched_requests"] = mgr.batched_requests
if not skip_cpu:
    t0 = time.perf_counter()
    want = R.workflow_fasta(Manager(R.reference_index()), mk(), sm, pre, msa)
    res["cpu_1core_s"] = time.perf_counter() - t0
    res["identical"] = bool(got == want)
    res["speedup"] = res["cpu_1core_s"] / res["gpu_s"]
res["note"] = "gpu_s: steady state (second run in the process); gpu_first_s: first full-size run; speedup = cpu_1core_s / gpu_s"
print(json.dumps(res))
