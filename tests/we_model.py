"""Host restatement of LocalMasterSlaveAligner + ProfileBuilder on top of the oracle
(praline/component/preprofile.py:227-267, util/align.py:187-266, container/align.py:30-61):
test infrastructure, pinned against tests/golden/local_ms.json (made by the reference itself)."""
import numpy as np

import oracle


def we_alignments(a, b, S, gaps, iterations):
    """The PairwiseAligner calls of one (master a, slave b): [(score, path, n_zero)] per iteration,
    zero_idxs growing by the bounding box of every path (preprofile.py:252-259)."""
    a, b = np.asarray(a), np.asarray(b)
    m = np.ascontiguousarray(np.asarray(S, np.float32)[a][:, b])
    g1, g2 = oracle.gap_arrays(len(a), len(b), gaps)
    zero, out = [], []
    for _ in range(iterations):
        score, path = oracle.align_raw("local", m, g1, g2, zero_idxs=zero if zero else None)
        out.append((score, path, len(zero)))
        ys, xs = path[:, 0], path[:, 1]
        zero.extend((y, x) for y in range(ys.min(), ys.max() + 1) for x in range(xs.min(), xs.max() + 1))
    return out


def boxes_of(paths):
    return [(int(p[:, 0].min()), int(p[:, 0].max()), int(p[:, 1].min()), int(p[:, 1].max())) for p in paths]


def compress_path(path):            # util/align.py:215-232, master column 0
    keep = [0] + [r for r in range(1, len(path)) if path[r, 0] > path[r - 1, 0]]
    return path[keep]


def extend_path_local(path, length):  # util/align.py:234-266, extend_idx 0
    first, last = path[0, 0], path[-1, 0]
    if first > 0:
        ext = np.full((first, 2), -1, int)
        ext[:, 0] = np.arange(first)
        path = np.vstack([ext, path])
    if last < length:
        ext = np.full((length - last, 2), -1, int)
        ext[:, 0] = np.arange(last + 1, length + 1)
        path = np.vstack([path, ext])
    return path


def local_master_counts(seqs, i, S, gaps, iterations, threshold, A):
    """ProfileBuilder's count table of master i against all other sequences, plus the merged path."""
    master = np.asarray(seqs[i])
    counts = np.zeros((len(master), A), np.int64)
    counts[np.arange(len(master)), master] += 1
    cols = [np.arange(len(master) + 1)]
    names = [i]
    for j, slave in enumerate(seqs):
        if j == i:
            continue
        slave = np.asarray(slave)
        for score, path, _ in we_alignments(master, slave, S, gaps, iterations):
            if threshold is not None and score < threshold:
                continue
            p = extend_path_local(compress_path(path.astype(int)), len(master))
            col = p[:, 1]
            cols.append(col)
            names.append(j)
            for r in range(len(master)):            # get_frequencies, util/align.py:205-211
                if col[r + 1] - col[r] > 0:
                    counts[r, slave[col[r + 1] - 1]] += 1
    return counts, np.stack(cols, axis=1), names
