"""GPU tree builder (rp_minmatch_*, relate_b200/csrc/minmatch.cu) against the oracle restatement of
MinMatch::QuickBuild (oracle/minmatch_oracle.c) and against the reference's own MinMatch (oracle/_ref/qblens):
the merge lists must be IDENTICAL, tree after tree on one handle (src/tree_builder.cpp:1060-1303, 2357-2646)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from relate_b200 import capi  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mm_cases  # noqa: E402

pytestmark = pytest.mark.gpu


def run_sequence(seed, N, kind, n_trees=4):
    o = oracle.MinMatchOracle(N, mm_cases.THETA)
    results = []

    def build(d, prior):
        m, info = o.quickbuild(d, prior)
        results.append((m, info))
        return m

    trees = mm_cases.tree_sequence(seed, N, kind, n_trees, oracle.prior_from_merges, build)
    with capi.MinMatch(N, mm_cases.THETA) as g:
        for t, (d, prior) in enumerate(trees):
            m, st = g.quickbuild(d, prior)
            om, info = results[t]
            first = int(np.argmax((m != om).any(axis=1))) if not np.array_equal(m, om) else -1
            assert first == -1, (f"{kind} N={N} seed={seed} tree {t}: first differing merge {first}: gpu {m[first]} "
                                 f"oracle {om[first]}; draws gpu {st['draws']} oracle {info['draws']}; fallback from "
                                 f"gpu {st['first_fallback_step']} oracle {info['first_sym_step']}")
            assert st["draws"] == info["draws"] and st["first_fallback_step"] == info["first_sym_step"]
    return trees, results


@pytest.mark.parametrize("kind", ["tree", "blocks", "uniform", "ties"])
@pytest.mark.parametrize("N", [2, 3, 5, 33, 100, 257])
def test_quickbuild_matches_oracle(kind, N):
    run_sequence(1000 + N, N, kind)


def test_quickbuild_n1000_matches_oracle_and_reference(tmp_path):
    N = 1000
    trees, results = run_sequence(7, N, "tree", n_trees=3)
    if os.access(oracle.REF_QBLENS, os.X_OK):
        ref, secs = oracle.reference_quickbuild(N, mm_cases.THETA, trees, str(tmp_path))
        for t in range(len(trees)):
            assert np.array_equal(ref[t], results[t][0]), f"oracle vs reference, tree {t}"


def test_small_pair_buffer_segments(monkeypatch):
    """RP_MINMATCH_CAP shrinks the pair buffer so that a tie-heavy Initialize needs several segments."""
    monkeypatch.setenv("RP_MINMATCH_CAP", "300")
    run_sequence(5, 130, "ties", n_trees=2)
    run_sequence(6, 130, "blocks", n_trees=3)


def test_handle_state_is_per_handle():
    """Two handles fed the same sequence give the same trees; a fresh handle fed only the last (d, prior) need not."""
    N = 64
    o = oracle.MinMatchOracle(N, mm_cases.THETA)
    trees = mm_cases.tree_sequence(3, N, "tree", 4, oracle.prior_from_merges, lambda d, p: o.quickbuild(d, p)[0])
    outs = []
    for _ in range(2):
        with capi.MinMatch(N, mm_cases.THETA) as g:
            outs.append([g.quickbuild(d, p)[0] for d, p in trees])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
