"""`praline` command line with the GPU aligners plugged in.

Runs the reference's own CLI (praline/cmd.py:24-137) unchanged; the only difference is the
manager it constructs for the single-process case: `Manager(index)` (cmd.py:43-44) becomes
`GpuBatchManager(index)`, which registers the GPU PairwiseAligner / RawPairwiseAligner under the
reference's type ids and batches the requests of each Execution.  INTEGRATION.md shows the
two-line patch that does the same inside praline/cmd.py behind a `--gpu` flag.

    python -m praline_b200.cli input.fasta output.aln --preprofile-global --msa-tree
"""
import sys


def main(argv=None):
    import praline.cmd as cmd
    from .plugin import GpuBatchManager
    cmd.Manager = GpuBatchManager
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    return cmd.main()


if __name__ == "__main__":
    main()
