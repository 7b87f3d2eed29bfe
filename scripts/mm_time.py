"""Times rp_minmatch_quickbuild on seeded GetMatrix-shaped matrices: python scripts/mm_time.py N [trees] [kind]"""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mm_cases
from relate_b200 import capi
from oracle import oracle
N = int(sys.argv[1]); T = int(sys.argv[2]) if len(sys.argv) > 2 else 3; kind = sys.argv[3] if len(sys.argv) > 3 else "tree"
with capi.MinMatch(N, mm_cases.THETA) as g:
    res = []; outs = []
    def build(d, prior):
        t0 = time.perf_counter(); m, st = g.quickbuild(d, prior); st["wall_ms"] = 1e3 * (time.perf_counter() - t0); res.append(st); return m
    trees = mm_cases.tree_sequence(1, N, kind, T, oracle.prior_from_merges, build)
for st in res: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
if os.access(oracle.REF_QBLENS, os.X_OK) and N <= 12000:
    with tempfile.TemporaryDirectory() as tmp:
        ref, secs = oracle.reference_quickbuild(N, mm_cases.THETA, trees, tmp)
    print(f"reference QuickBuild: {1e3 * secs / T:.2f} ms/tree (CPU, 1 core); merge lists identical: {all(np.array_equal(a, b) for a, b in zip(ref, outs))}")
