"""End-to-end MSA wall time (BASELINE metric, second half): the reference's
PralineMultipleSequenceAlignmentWorkflow (what `praline in.fa out.aln --preprofile-global
--msa-tree` runs, praline/cmd.py:56-119) on ONE manager per process:

    python tools/msa_e2e.py <n_seqs> <length> <preprofile> <msa> <arm>

arm = gpu   GpuBatchManager (first full-size run and steady state of the same process)
      cpu1  the reference's stock Manager, one core            (`praline -t 1`,  cmd.py:41-44)
      cpun  the reference's ParallelExecutionManager, all host cores (`praline -t $(nproc)`,
            cmd.py:41-42, core/manager.py:282-477: nproc - 1 forked workers, as the CLI does)

Prints one JSON line with the wall time(s) and the SHA-256 of the FASTA output; the caller
(bench.py) compares the hashes of the arms -- outputs must be byte-identical.  Every arm is its
own process so that the forked workers of `cpun` never see a CUDA context.
"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_praline as R
from praline_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pre = sys.argv[3] if len(sys.argv) > 3 else "global"
msa = sys.argv[4] if len(sys.argv) > 4 else "tree"
arm = sys.argv[5] if len(sys.argv) > 5 else "gpu"

import praline
from praline.core import Manager, ParallelExecutionManager
from praline.container import Sequence, PlainTrack, ALPHABET_AA, TRACK_ID_INPUT

with praline.open_builtin('matrices/blosum62') as f:
    sm = praline.load_score_matrix(f, alphabet=ALPHABET_AA)
fam = synth.family(1, n, L)
mk = lambda: [Sequence("s%d" % i, [(TRACK_ID_INPUT, PlainTrack(None, ALPHABET_AA, raw_indices=s))]) for i, s in enumerate(fam)]
sha = lambda text: hashlib.sha256(text.encode()).hexdigest()

res = {"n_seqs": n, "length": L, "preprofile": pre, "msa": msa, "arm": arm, "cores": os.cpu_count()}
if arm == "gpu":
    from praline_b200 import plugin
    mgr = plugin.GpuBatchManager(R.reference_index())
    R.workflow_fasta(mgr, mk()[:4], sm, pre, msa)          # warm-up: CUDA context, kernels
    t0 = time.perf_counter()
    got = R.workflow_fasta(mgr, mk(), sm, pre, msa)
    res["first_s"] = time.perf_counter() - t0             # first full-size run of the process (allocator, pinned staging warm up)
    t0 = time.perf_counter()
    got2 = R.workflow_fasta(mgr, mk(), sm, pre, msa)
    res["wall_s"] = time.perf_counter() - t0              # steady state (second run in the process)
    assert got2 == got
    res["batched_requests"] = mgr.batched_requests
elif arm == "cpu1":
    t0 = time.perf_counter()
    got = R.workflow_fasta(Manager(R.reference_index()), mk(), sm, pre, msa)
    res["wall_s"] = time.perf_counter() - t0
    res["threads"] = 1
elif arm == "cpun":
    cores = os.cpu_count() or 1
    mgr = ParallelExecutionManager(R.reference_index(), max(cores - 1, 1))     # cmd.py:42: num_threads - 1 workers
    t0 = time.perf_counter()
    got = R.workflow_fasta(mgr, mk(), sm, pre, msa)
    res["wall_s"] = time.perf_counter() - t0
    res["threads"] = cores
    try:
        mgr.close()
    except Exception:
        pass
else:
    raise SystemExit("unknown arm " + arm)
res["sha256"] = sha(got)
res["fasta_bytes"] = len(got)
print(json.dumps(res))
