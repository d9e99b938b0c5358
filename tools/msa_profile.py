import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_praline as R
from praline_b200 import synth, plugin
import praline
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT
pre, msa = sys.argv[1], sys.argv[2]
with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
fam = synth.family(1, int(sys.argv[3]) if len(sys.argv) > 3 else 50, 300)
mk = lambda: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=s))]) for i, s in enumerate(fam)]
mgr = plugin.GpuBatchManager(R.reference_index())
R.workflow_fasta(mgr, mk()[:4], sm, pre, msa)
pr = cProfile.Profile(); pr.enable()
R.workflow_fasta(mgr, mk(), sm, pre, msa)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
