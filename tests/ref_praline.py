"""Makes the reference package importable in tests (baseline/_ref, installed from
/root/reference with `pip install --no-deps --target baseline/_ref`; git-ignored, it travels
to the GPU box).  HAVE_PRALINE is False when it is absent -- those tests then skip."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
HAVE_PRALINE = False
if os.path.isdir(os.path.join(_REF, "praline")):
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    try:
        import itsdangerous  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "tests", "_stubs"))
    warnings.filterwarnings("ignore")
    try:
        import praline  # noqa: F401
        import praline.component  # noqa: F401
        HAVE_PRALINE = True
    except Exception:  # pragma: no cover
        HAVE_PRALINE = False

ROOT_TAG = "__ROOT_TAG__"


def reference_index():
    """TypeIndex with the reference's components registered by hand (the package is not
    installed with entry points on every box)."""
    from praline.core import TypeIndex
    import praline.component as pc
    index = TypeIndex()
    for name in ("PairwiseAligner", "RawPairwiseAligner", "DummyMasterSlaveAligner", "GlobalMasterSlaveAligner",
                 "LocalMasterSlaveAligner", "AdHocMultipleSequenceAligner", "TreeMultipleSequenceAligner",
                 "ProfileBuilder", "GuideTreeBuilder", "PralineMultipleSequenceAlignmentWorkflow"):
        index.register(getattr(pc, name))
    return index


def run_task(manager, component, env_keys, **inputs):
    from praline.core import Execution, Environment
    ex = Execution(manager, ROOT_TAG)
    task = ex.add_task(component)
    task.environment(Environment(keys=env_keys))
    task.inputs(**inputs)
    msgs = [m for m in ex.run()]
    return ex.outputs[0], msgs


def workflow_fasta(manager, seqs, score_matrix, preprofile="global", msa="tree", gaps=(-11.0, -1.0), extra=None):
    """Run the reference's MSA workflow (what `praline --preprofile-global --msa-tree` runs,
    praline/cmd.py:56-119) on the given manager; returns the FASTA text of the alignment."""
    import io
    import praline
    import praline.component as pc
    from praline.container import TRACK_ID_INPUT
    keys = {'gap_series': [float(g) for g in gaps], 'linkage_method': 'average', 'aligner': pc.PairwiseAligner.tid,
            'merge_mode': 'global', 'dist_mode': 'global', 'msa_mode': msa, 'preprofile_mode': preprofile,
            'score_threshold': None, 'waterman_eggert_iterations': 2, 'debug': 0, 'accelerate': True}
    if extra:
        keys.update(extra)
    out, _ = run_task(manager, pc.PralineMultipleSequenceAlignmentWorkflow, keys, sequences=seqs,
                      score_matrix=score_matrix)
    import tempfile
    with tempfile.NamedTemporaryFile("r", suffix=".fa") as f:
        praline.write_alignment_fasta(f.name, out['alignment'], TRACK_ID_INPUT)
        return open(f.name).read()
