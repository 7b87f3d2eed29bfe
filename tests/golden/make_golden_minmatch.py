"""Writes tests/golden/minmatch/merges_ref.json: merge lists produced by the REFERENCE's MinMatch::QuickBuild
(oracle/_ref/qblens, compiled from /root/reference) for the seeded matrix sequences of tests/mm_cases.py.
Run in the build container:  python tests/golden/make_golden_minmatch.py"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
import mm_cases  # noqa: E402

CASES = [(31, 5, "uniform"), (32, 24, "ties"), (33, 48, "tree"), (34, 40, "blocks"), (35, 33, "uniform")]
out = {"theta": mm_cases.THETA, "generator": "oracle/_ref/qblens on tests/mm_cases.tree_sequence(seed, N, kind, 4)", "cases": []}
for seed, N, kind in CASES:
    # the next prior needs the previous tree: build the sequence one tree at a time with the reference itself
    trees, merges = [], []

    def build(d, prior):
        trees.append((d, prior))
        with tempfile.TemporaryDirectory() as tmp:
            ref, _ = oracle.reference_quickbuild(N, mm_cases.THETA, trees, tmp)
        merges.append(ref[-1])
        return ref[-1]

    mm_cases.tree_sequence(seed, N, kind, 4, oracle.prior_from_merges, build)
    out["cases"].append({"seed": seed, "N": N, "kind": kind, "merges": [m.tolist() for m in merges]})
json.dump(out, open(os.path.join(HERE, "minmatch", "merges_ref.json"), "w"))
print("wrote", len(out["cases"]), "cases")
