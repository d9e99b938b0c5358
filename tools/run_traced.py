"""Small driver used under ncu: one traced pass (K2 with traceback + K4) over 60k pairs of the
bench workload, and one general-kernel alignment; prints device times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from praline_b200 import get_engine, matrices, synth

eng = get_engine(0)
S = matrices.blosum62()
seqs = synth.family(2, 1000, 300)
batch = eng.batch(seqs)
pi, pj = synth.all_pairs(len(seqs))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
pi, pj = pi[:n], pj[:n]
cells = int((batch.lens[pi] * batch.lens[pj]).sum())
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.align_pairs(batch, pi, pj, S, [-11.0, -1.0], mode="global", want_paths=True, resident="one", device_only=True)
    b.record()
    torch.cuda.synchronize()
    print("traced pass %d: %.3f ms, %.1f GCUPS" % (rep, a.elapsed_time(b), cells / a.elapsed_time(b) / 1e6))
