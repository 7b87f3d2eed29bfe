import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, mm_cases
from relate_b200 import capi
from oracle import oracle
for seed, N, kind in [(1, 100, "tree"), (2, 70, "ties"), (3, 90, "blocks"), (4, 60, "uniform")]:
    o = oracle.MinMatchOracle(N, mm_cases.THETA); res = []
    trees = mm_cases.tree_sequence(seed, N, kind, 2, oracle.prior_from_merges, lambda d, p: res.append(o.quickbuild(d, p)[0]) or res[-1])
    with capi.MinMatch(N, mm_cases.THETA) as g:
        for t, (d, p) in enumerate(trees):
            m, st = g.quickbuild(d, p)
            assert np.array_equal(m, res[t]), (kind, t)
    print(kind, "ok", st)
