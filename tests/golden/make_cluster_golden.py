#!/usr/bin/env python
"""Golden merge orders for the guide-tree clustering, made by RUNNING THE REFERENCE's
HierarchicalClusteringAlgorithm (praline/util/cluster.py:15-57) on seeded distance matrices:
integer-valued ones with many ties (the sequence-score case, where the order is exact) and
f32 ones.  Needs the reference in baseline/_ref (build container only).

    python tests/golden/make_cluster_golden.py   ->  tests/golden/cluster.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests", "_stubs"))
import warnings
warnings.filterwarnings("ignore")
from praline.util import HierarchicalClusteringAlgorithm  # noqa: E402


def sym(rng, n, kind):
    if kind == "ties":          # small integer range: many exact ties
        s = rng.integers(-20, 21, (n, n)).astype(np.float32)
    elif kind == "int":         # score-like integers
        s = rng.integers(-400, 1500, (n, n)).astype(np.float32)
    else:                       # f32 profile-score-like values
        s = (rng.standard_normal((n, n)) * 37.0).astype(np.float32)
    s = np.triu(s, 1)
    d = s + s.T                 # diagonal 0, as GuideTreeBuilder builds it (tree.py:92-133)
    return ((-d) + d.max()).astype(np.float32)


cases = []
rng = np.random.default_rng(11)
for n, kind in [(2, "int"), (3, "ties"), (5, "ties"), (8, "ties"), (13, "int"), (21, "ties"), (34, "int"),
                (34, "f32"), (48, "ties"), (60, "f32")]:
    dist = sym(rng, n, kind)
    for linkage in ("single", "complete", "average"):
        order = list(HierarchicalClusteringAlgorithm(dist).merge_order(linkage))
        cases.append({"n": n, "kind": kind, "linkage": linkage, "dist": dist.tolist(),
                      "order": [[int(a), int(b)] for a, b in order]})
with open(os.path.join(HERE, "cluster.json"), "w") as f:
    json.dump(cases, f)
print("wrote %d cases" % len(cases))
