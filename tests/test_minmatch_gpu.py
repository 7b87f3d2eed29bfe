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


@pytest.mark.parametrize("env", [{"RP_MINMATCH_GENERAL": "1"}, {"RP_MINMATCH_GENERAL": "2"}, {"RP_MINMATCH_NO_SMEM": "1"},
                                 {"RP_MINMATCH_THREADS": "256"}, {"RP_MINMATCH_THREADS": "1024"},
                                 {"RP_MINMATCH_NO_SMEM": "1", "RP_MINMATCH_GENERAL": "2", "RP_MINMATCH_THREADS": "1024"}])
def test_every_code_path_gives_the_same_trees(monkeypatch, env):
    """The kernel picks a path per merge step by size (small / medium / any-size pair handling, block-wide or warp-per-row
    rescans), keeps the per-cluster arrays in shared or in global memory by N, and runs with 256, 512 or 1024 threads: each
    choice forced in turn on inputs that normally take another one."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for seed, N, kind in [(41, 130, "tree"), (42, 257, "blocks"), (43, 100, "ties"), (44, 64, "uniform"), (45, 600, "tree")]:
        run_sequence(seed, N, kind, n_trees=3)


def test_matrices_already_on_the_device():
    """rp_minmatch_quickbuild_device: the matrices are taken from device memory (as the distance kernel leaves them) and are
    not modified; same trees as from the host."""
    import torch
    N = 200
    o = oracle.MinMatchOracle(N, mm_cases.THETA)
    res = []
    trees = mm_cases.tree_sequence(9, N, "tree", 3, oracle.prior_from_merges, lambda d, p: res.append(o.quickbuild(d, p)[0]) or res[-1])
    with capi.MinMatch(N, mm_cases.THETA) as g:
        for t, (d, prior) in enumerate(trees):
            dd = torch.from_numpy(d).cuda()
            pp = torch.from_numpy(prior).cuda() if prior is not None else None
            torch.cuda.synchronize()
            m, _ = g.quickbuild_device(dd.data_ptr(), pp.data_ptr() if pp is not None else None)
            assert np.array_equal(m, res[t])
            assert np.array_equal(dd.cpu().numpy(), d) and (pp is None or np.array_equal(pp.cpu().numpy(), prior))


def test_handles_of_different_sizes_coexist():
    """The opt-in shared-memory limit is a per-kernel setting shared by all handles: a small handle created after a large one
    must not take the large one's shared memory away."""
    big, small = 700, 520   # both run the 512-thread instance; 28 KB vs 21 KB of per-cluster state
    rng = np.random.default_rng(3)
    db, ds = mm_cases.matrix(rng, big, "tree"), mm_cases.matrix(rng, small, "tree")
    want_b, _ = oracle.MinMatchOracle(big, mm_cases.THETA).quickbuild(db)
    want_s, _ = oracle.MinMatchOracle(small, mm_cases.THETA).quickbuild(ds)
    with capi.MinMatch(big, mm_cases.THETA) as gb:
        with capi.MinMatch(small, mm_cases.THETA) as gs:
            assert np.array_equal(gs.quickbuild(ds)[0], want_s)
            assert np.array_equal(gb.quickbuild(db)[0], want_b)
        gb2 = capi.MinMatch(big, mm_cases.THETA)
        assert np.array_equal(gb2.quickbuild(db)[0], want_b)
        gb2.close()


def test_reset_gives_a_fresh_object():
    """rp_minmatch_reset: the same tree sequence before and after a reset gives the same trees (state carried from tree to tree
    would change the second pass otherwise: the sequence is NOT idempotent without the reset)."""
    N = 96
    o = oracle.MinMatchOracle(N, mm_cases.THETA)
    trees = mm_cases.tree_sequence(12, N, "blocks", 4, oracle.prior_from_merges, lambda d, p: o.quickbuild(d, p)[0])
    with capi.MinMatch(N, mm_cases.THETA) as g:
        first = [g.quickbuild(d, p)[0] for d, p in trees]
        g.reset()
        second = [g.quickbuild(d, p)[0] for d, p in trees]
    for a, b in zip(first, second):
        assert np.array_equal(a, b)


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


# ---- through the reference's BuildTopology (oracle/_ref/Relate_gpu: GetMatrix and QuickBuild bound to the C ABI) ------
import filecmp  # noqa: E402
import hashlib  # noqa: E402
import json  # noqa: E402
import re  # noqa: E402
import subprocess  # noqa: E402
import time  # noqa: E402

from relate_b200 import synth  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
need_relate_gpu = pytest.mark.skipif(not os.access(oracle.REF_RELATE_GPU, os.X_OK), reason="oracle/_ref/Relate_gpu did not travel")


def _bt(cwd, W, out="o", painting="0.001,1", seed="1", **env):
    t0 = time.perf_counter()
    p = subprocess.run([oracle.REF_RELATE_GPU, "--mode", "BuildTopology", "--chunk_index", "0", "--first_section", "0",
                        "--last_section", str(W - 1), "-o", out, "--painting", painting, "--seed", seed],
                       cwd=cwd, capture_output=True, text=True,
                       env=dict(os.environ, RELATE_GPU_MINMATCH_STATS="1", RELATE_GPU_MINMATCH_MIN_N="0", **env))
    assert p.returncode == 0, p.stderr[-2000:]
    m = re.search(r"QuickBuild: (\d+) trees on the GPU \((\d+) by the reference's code\), ([0-9.]+) s in the call, ([0-9.]+) s in the kernel",
                  p.stderr)
    assert m, p.stderr[-2000:]
    return dict(wall=time.perf_counter() - t0, gpu_trees=int(m.group(1)), ref_trees=int(m.group(2)), call_s=float(m.group(3)),
                kernel_s=float(m.group(4)))


@need_relate_gpu
def test_buildtopology_trees_on_gpu_example_md5(tmp_path):
    """Bundled example: GPU Paint -> BuildTopology with GPU d_ij AND GPU trees writes the reference's .anc/.mut (golden md5s);
    with the built-in cross-check every tree is also compared merge by merge with the reference's QuickBuild."""
    from conftest import unpack_golden
    unpack_golden("example_c1", str(tmp_path), out="ex")
    meta = json.load(open(os.path.join(GOLDEN, "example_c1", "topology_md5.json")))
    capi.paint_chunk(str(tmp_path / "ex"), 0, meta["painting"])
    for env in ({}, {"RELATE_GPU_MINMATCH_VERIFY": "1"}):
        st = _bt(str(tmp_path), meta["W"], out="ex", painting=meta["painting"], seed=str(meta["seed"]), **env)
        assert st["gpu_trees"] > 0 and st["ref_trees"] == 0
        for fn, want in meta["md5"].items():
            got = hashlib.md5(open(os.path.join(str(tmp_path), "ex", "chunk_0", fn), "rb").read()).hexdigest()
            assert got == want, (fn, env)


@need_relate_gpu
@pytest.mark.parametrize("N,L,W", [(200, 3000, 3), (1000, 3000, 1)])
def test_buildtopology_gpu_trees_byte_identical_to_cpu_trees(tmp_path, N, L, W):
    """Same paint files, same GPU distance matrices; trees by the reference's CPU QuickBuild (RELATE_GPU_MINMATCH=0) vs by
    rp_minmatch_quickbuild: the .anc/.mut files must be byte-identical.  Prints what the tree builder costs each way."""
    for tag in ("cpu", "gpu"):
        synth.make_chunk_dir(str(tmp_path / tag / "o"), N, L, seed=31, n_windows=W)
        capi.paint_chunk(str(tmp_path / tag / "o"), 0, "0.001,1")
    a = _bt(str(tmp_path / "cpu"), W, RELATE_GPU_MINMATCH="0")
    b = _bt(str(tmp_path / "gpu"), W)
    assert a["gpu_trees"] == 0 and b["ref_trees"] == 0 and a["ref_trees"] == b["gpu_trees"] > 0
    for w in range(W):
        for ext in ("anc", "mut"):
            assert filecmp.cmp(str(tmp_path / "cpu" / "o" / "chunk_0" / f"o_{w}.{ext}"),
                               str(tmp_path / "gpu" / "o" / "chunk_0" / f"o_{w}.{ext}"), shallow=False), (w, ext)
    print(f"\nBuildTopology N={N} L={L}: {b['gpu_trees']} trees; QuickBuild on the CPU {a['call_s']:.3f} s "
          f"({1e3 * a['call_s'] / a['ref_trees']:.2f} ms/tree), on the GPU {b['call_s']:.3f} s in the call, {b['kernel_s']:.3f} s in the "
          f"kernel ({1e3 * b['kernel_s'] / b['gpu_trees']:.2f} ms/tree); BuildTopology wall {a['wall']:.2f} s -> {b['wall']:.2f} s")


@need_relate_gpu
def test_cli_gpu_topology_flag(tmp_path):
    """`relate --mode BuildTopology --gpu_topology` = BuildTopology in Relate_gpu (GPU window repaint, distance matrices and trees)
    on the paint files `relate --mode Paint` wrote: the same files as running Relate_gpu by hand."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "relate_b200", "bin", "relate")
    N, L, W = 200, 2000, 2
    for tag in ("cli", "direct"):
        synth.make_chunk_dir(str(tmp_path / tag / "o"), N, L, seed=33, n_windows=W)
    env = dict(os.environ, RELATE_REFERENCE_BIN=oracle.REF_RELATE, RELATE_GPU_BIN=oracle.REF_RELATE_GPU, RELATE_GPU_MINMATCH_MIN_N="0")
    bt = ["--mode", "BuildTopology", "--chunk_index", "0", "--first_section", "0", "--last_section", str(W - 1), "-o", "o",
          "--painting", "0.001,1", "--seed", "1"]
    for tag, cmd in (("cli", [exe] + bt + ["--gpu_topology"]), ("direct", [oracle.REF_RELATE_GPU] + bt)):
        p = subprocess.run([exe, "--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", "0.001,1"], cwd=str(tmp_path / tag),
                           capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr[-1000:]
        p = subprocess.run(cmd, cwd=str(tmp_path / tag), capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr[-1000:]
    for w in range(W):
        for ext in ("anc", "mut"):
            assert filecmp.cmp(str(tmp_path / "cli" / "o" / "chunk_0" / f"o_{w}.{ext}"),
                               str(tmp_path / "direct" / "o" / "chunk_0" / f"o_{w}.{ext}"), shallow=False), (w, ext)
