"""Seeded synthetic inputs of the benchmark shapes.

Family generator: a root of length L drawn uniformly over the first `n_sym`
alphabet indices (20 standard residues / 4 nucleotides); every member is the
root with 30 % i.i.d. substitutions and 3 indels of length 1-5, so that
alignments are non-degenerate (unrelated random sequences give all-end-gap
semiglobal alignments).
"""
import numpy as np


def family(seed, n, length, n_sym=20, sub_rate=0.30, n_indels=3, max_indel=5):
    """Returns a list of n int32 index arrays."""
    rng = np.random.default_rng(seed)
    root = rng.integers(0, n_sym, size=length, dtype=np.int32)
    out = []
    for _ in range(n):
        s = root.copy()
        mask = rng.random(length) < sub_rate
        s[mask] = rng.integers(0, n_sym, size=int(mask.sum()), dtype=np.int32)
        for _ in range(n_indels):
            k = int(rng.integers(1, max_indel + 1))
            pos = int(rng.integers(0, len(s) + 1))
            if rng.random() < 0.5 and len(s) > k + 1:
                pos = min(pos, len(s) - k)
                s = np.delete(s, np.arange(pos, pos + k))
            else:
                s = np.insert(s, pos, rng.integers(0, n_sym, size=k, dtype=np.int32))
        out.append(np.ascontiguousarray(s, dtype=np.int32))
    return out


def pack(seqs):
    """list of index arrays -> (flat int32, int64 offsets[n+1])."""
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    flat = np.concatenate(seqs).astype(np.int32) if len(seqs) else np.zeros(0, np.int32)
    return flat, offs


def all_pairs(n):
    """Unordered pairs (i < j) in the order GuideTreeBuilder queues them
    (reference: praline/component/tree.py:99-131)."""
    i, j = np.triu_indices(n, k=1)
    return i.astype(np.int32), j.astype(np.int32)


def count_profile(seed, length, depth, n_sym, alphabet_size):
    """A depth-`depth` count profile [length x alphabet_size] (int64), i.e. what a
    ProfileTrack holds (reference: praline/container/sequence.py:176-203)."""
    fam = family(seed, depth, length, n_sym=n_sym, n_indels=0)
    counts = np.zeros((length, alphabet_size), np.int64)
    for s in fam:
        counts[np.arange(length), s[:length]] += 1
    return counts


def profile_from_counts(counts):
    """counts -> f32 probabilities exactly as ProfileTrack.profile does
    (reference: praline/container/sequence.py:198-202: f32 totals, f64 divide, f32 cast)."""
    totals = np.array(counts.sum(axis=1), dtype=np.float32)
    return np.array(counts / totals[:, np.newaxis], dtype=np.float32)
