"""Distance matrices for the tree-builder parity tests (shared by the CPU and the GPU tests)."""
import numpy as np

THETA = float(np.float32(0.001))
VAL = float(-np.log(THETA / (1 - THETA)))


def matrix(rng, N, kind):
    """uniform: no ties, the symmetric fallback kicks in early; ties: a handful of distinct values, thousands of draws
    per step; tree: GetMatrix-shaped (mutation counts below the MRCA on a random coalescent tree times log-odds, a little
    noise, row minimum subtracted, zero diagonal); blocks: groups of identical haplotypes (exact zeros off the diagonal)."""
    if kind == "uniform":
        d = rng.random((N, N), dtype=np.float32) * 20
    elif kind == "ties":
        d = rng.integers(0, 4, (N, N)).astype(np.float32) * np.float32(VAL)
    elif kind in ("tree", "blocks"):
        depth = np.zeros(N)
        d = np.zeros((N, N))
        mem = {k: [k] for k in range(N)}
        alive = list(range(N))
        rate = 1.5 if kind == "tree" else 0.4
        while len(alive) > 1:
            a, b = rng.choice(len(alive), 2, replace=False)
            a, b = alive[a], alive[b]
            for side in (a, b):
                add = rng.poisson(rate)
                for x in mem[side]:
                    depth[x] += add
            ia, ib = np.array(mem[a]), np.array(mem[b])
            d[np.ix_(ia, ib)] = depth[ia][:, None]
            d[np.ix_(ib, ia)] = depth[ib][:, None]
            mem[a] = mem[a] + mem[b]
            del mem[b]
            alive.remove(b)
        d = d * VAL
        if kind == "tree":
            d = d + (rng.random((N, N)) < 0.3) * rng.random((N, N)) * 0.5
        d = d.astype(np.float32)
        np.fill_diagonal(d, np.inf)
        d -= d.min(axis=1, keepdims=True)
    else:
        raise ValueError(kind)
    d = d.astype(np.float32)
    np.fill_diagonal(d, 0)
    return d


def tree_sequence(seed, N, kind, n_trees=4, prior_from_merges=None, build=None):
    """A list of (d, prior) as BuildTopology feeds them to one MinMatch object: the first tree without a prior, the later
    ones with the prior derived from the previous tree; the last with a prior that no tree explains.  `build(d, prior)`
    returns the merges of the tree just built (needed for the next prior)."""
    rng = np.random.default_rng(seed)
    trees, prev = [], None
    for t in range(n_trees):
        d = matrix(rng, N, kind)
        if t == 0:
            prior = None
        elif t == n_trees - 1 and n_trees > 2:
            prior = matrix(rng, N, "ties")
        else:
            prior = prior_from_merges(prev, N, VAL)
        trees.append((d, prior))
        prev = build(d, prior)
    return trees
