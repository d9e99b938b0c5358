"""MSA wall time at 200 and 500 sequences (300 aa, `--preprofile-global --msa-tree`): GpuBatchManager against the
reference's ParallelExecutionManager with all host cores (`-t $(nproc)`), outputs compared by SHA-256.
Writes profiles-style JSON to stdout:  python tools/msa_full.py > gpurun_out/r02_msa_e2e_full.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
out = {}
for n in (200, 500):
    out["tree_%d" % n] = bench.msa_e2e(n, 300, "global", "tree", arms=("gpu", "cpun"))
out["cli_default_200"] = bench.msa_e2e(200, 300, "dummy", "ad_hoc", arms=("gpu", "cpun"))
print(json.dumps(out))
