"""cProfile of the timed MSA workflow run of tools/msa_e2e.py (GpuBatchManager only)."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ref_praline as R
from praline_b200 import synth, plugin
import praline
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT
n, L = int(sys.argv[1]), int(sys.argv[2])
pre, msa = sys.argv[3], sys.argv[4]
with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
fam = synth.family(1, n, L)
mk = lambda: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=s))]) for i, s in enumerate(fam)]
mgr = plugin.GpuBatchManager(R.reference_index())
R.workflow_fasta(mgr, mk()[:4], sm, pre, msa)
for rep in range(3):
    t0 = time.perf_counter(); R.workflow_fasta(mgr, mk(), sm, pre, msa); print("run %d: %.3f s" % (rep, time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
R.workflow_fasta(mgr, mk(), sm, pre, msa)
pr.disable()
st = pstats.Stats(pr).sort_stats("tottime")
st.print_stats(8)
st.print_callers("torch.empty")
import torch
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
