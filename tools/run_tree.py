"""Guide-tree stage on the device at scale: all-vs-all scores -> distance matrix -> clustering
kernel (csrc/cluster.cu), with the reference's pure-Python clustering (util/cluster.py) timed on
a small subsample for comparison (it is O(n^3) Python; 10^4 sequences are out of its reach).

    python tools/run_tree.py [n_seqs=2000] [length=400] [ref_n=60]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
ref_n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
eng = get_engine(0)
S = matrices.blosum62()
seqs = synth.family(4, n, L)
batch = eng.batch(seqs)
S_dev = eng.dev(S)
eng.cluster_merge_order(np.zeros((4, 4), np.float32))      # warm-up
for rep in range(2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record()
    cond, _, cells = eng.allpairs_scores(batch, S_dev, 27, [-11.0, -1.0], mode="global", S_host=S)
    ev[1].record()
    dist = eng.tree_distance(cond, n)
    ev[2].record()
    t0 = time.perf_counter()
    merges = eng.cluster_merge_order(dist, "average")
    wall = time.perf_counter() - t0
    ev[3].record()
    torch.cuda.synchronize()
res = {"stage": "guide tree on device", "n_seqs": n, "length": L, "pairs": n * (n - 1) // 2,
       "allpairs_ms": ev[0].elapsed_time(ev[1]), "distance_ms": ev[1].elapsed_time(ev[2]),
       "cluster_ms": ev[2].elapsed_time(ev[3]), "cluster_wall_ms_incl_d2h": 1e3 * wall,
       "gcups_allpairs": cells / ev[0].elapsed_time(ev[1]) / 1e6, "merges": len(merges), "first_merges": merges[:3]}
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import ref_praline as R
    if R.HAVE_PRALINE and ref_n > 1:
        from praline.util import HierarchicalClusteringAlgorithm
        sub = dist[:ref_n, :ref_n].cpu().numpy()
        t0 = time.perf_counter()
        want = list(HierarchicalClusteringAlgorithm(sub).merge_order("average"))
        res["reference_cluster_s_at_n%d" % ref_n] = time.perf_counter() - t0
        res["reference_same_order_at_n%d" % ref_n] = [tuple(int(v) for v in w) for w in want] == eng.cluster_merge_order(sub, "average")
except Exception as e:   # the device numbers stand on their own
    res["reference_error"] = str(e)[:200]
print(json.dumps(res))
