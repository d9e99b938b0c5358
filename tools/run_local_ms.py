"""Local master-slave preprofile stage (LocalMasterSlaveAligner x N + ProfileBuilder,
preprofile.py:160-267) on the device: N sequences, all ordered pairs, Waterman-Eggert iterations;
prints device + host time and GCUPS (cells = pairs x L1 x L2 x iterations).  With `wf` also times
the reference's workflow (preprofile_mode='local', tree MSA) on GpuBatchManager vs the stock Manager.

    python tools/run_local_ms.py [n_seqs] [length] [iterations] [wf]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
length = int(sys.argv[2]) if len(sys.argv) > 2 else 300
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = get_engine(0)
S = matrices.blosum62()
seqs = synth.family(2, n, length)
batch = eng.batch(seqs)
masters = np.repeat(np.arange(n), n - 1)
slaves = np.concatenate([np.delete(np.arange(n), i) for i in range(n)])
cells = int((batch.lens[masters] * batch.lens[slaves]).sum()) * iters
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cnt, where, scores = eng.local_preprofile_counts(batch, masters, slaves, S, [-11.0, -1.0], iterations=iters)
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("local preprofiles %d x %d, %d iterations, pass %d: device %.2f ms, wall %.1f ms, %.1f GCUPS (wall)"
          % (n, length, iters, rep, a.elapsed_time(b), wall * 1e3, cells / wall / 1e9))
eng.dump_trace()
print("mean score per iteration:", scores.mean(axis=1))

if "wf" in sys.argv:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import ref_praline as R
    import praline
    from praline.core import Manager
    from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT
    from praline_b200 import plugin
    with praline.open_builtin('matrices/blosum62') as f:
        sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
    fam = synth.family(1, 30, 200)
    mk = lambda: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=np.asarray(s)))])
                  for i, s in enumerate(fam)]
    t0 = time.perf_counter()
    got = R.workflow_fasta(plugin.GpuBatchManager(R.reference_index()), mk(), sm, "local", "tree")
    t1 = time.perf_counter()
    got = R.workflow_fasta(plugin.GpuBatchManager(R.reference_index()), mk(), sm, "local", "tree")
    t2 = time.perf_counter()
    want = R.workflow_fasta(Manager(R.reference_index()), mk(), sm, "local", "tree")
    t3 = time.perf_counter()
    print("workflow local/tree 30 x 200: GPU manager %.3f s (first %.3f s), stock Manager %.2f s, identical: %s"
          % (t2 - t1, t1 - t0, t3 - t2, got == want))
