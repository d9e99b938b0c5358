"""cProfile of the C4 workflow on GpuBatchManager (where does the wall time go)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_praline as R
from praline_b200 import synth, plugin
import praline
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
fam = synth.family(4, n, 400)
mk = lambda k: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=s))]) for i, s in enumerate(fam[:k])]
mgr = plugin.GpuBatchManager(R.reference_index())
extra = {'merge_mode': 'semiglobal'}
R.workflow_fasta(mgr, mk(4), sm, "global", "tree", extra=extra)
pr = cProfile.Profile(); pr.enable()
R.workflow_fasta(mgr, mk(n), sm, "global", "tree", extra=extra)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40); print(s.getvalue()[:8000])
