"""CPU tests of the oracle (oracle/paint_oracle.c) against the reference's golden vectors.

The oracle is test infrastructure; these tests are what "pins" it:
  * cmp-identical paint files vs fixtures the unmodified reference binary produced (tests/golden/),
  * the reference's own known-answer material for the path: the 5x5 mismatch-count matrix of
    include/test/test_painting.cpp:54-81 and the fast_log tolerance of include/test/test_log.cpp:5-15,
  * a fresh differential run against oracle/_ref/Relate when that binary is present.
"""
import filecmp
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, make_case, unpack_golden
from oracle import oracle
from relate_b200 import chunkio, synth


@pytest.mark.parametrize("name,painting,ref_dir", [
    ("example_c1", "0.001,1", "paint_ref"),
    ("synth_n96", "0.001,1", "paint_ref"),
    ("synth_n96", None, "paint_ref_noflag"),
])
def test_oracle_reproduces_reference_paint_files(tmp_path, name, painting, ref_dir):
    d = unpack_golden(name, str(tmp_path))
    st = oracle.paint_chunk(d, 0, painting)
    for w in range(st["W"]):
        mine = os.path.join(d, "chunk_0", "paint", f"relate_{w}.bin")
        ref = os.path.join(GOLDEN, name, ref_dir, f"relate_{w}.bin")
        assert filecmp.cmp(mine, ref, shallow=False), f"window {w} differs from the reference's bytes"


def test_fast_log_tolerance_like_reference_test_log():
    # include/test/test_log.cpp:5-15
    x = np.float32(1e-5) * np.float32(10.0) ** np.arange(1, 10, dtype=np.float32)
    assert np.all(np.abs(oracle.fast_log(x) - np.log(x.astype(np.float64))) < 0.007)


# data of include/test/test_painting.cpp:33-52 (N=5, L=10) and its hard-coded matrix d (:54-81)
KAT_ROWS = ["0110000000", "0110010100", "0100000000", "0000100000", "0000100000"]
KAT_D = np.array([[0, 0, 1, 2, 2], [2, 0, 3, 4, 4], [0, 0, 0, 1, 1], [1, 1, 1, 0, 0], [1, 1, 1, 0, 0]])


def kat_inputs():
    N, L = 5, 10
    hap = np.zeros((L, N), np.uint8)
    for n, row in enumerate(KAT_ROWS):
        hap[:, n] = np.frombuffer(row.encode(), np.uint8)
    r = np.zeros(L)
    wb = np.array([0, 5, L], np.int32)
    return hap, r, wb, 0.025


def kat_check(res, hap, theta):
    """With r=0, alpha at visited site i times beta at visited site i+1 is (ntheta/(N-1)) * tau^(d - mis_{i+1}):
    window 1's alpha sits at the last k-site < 5 and window 0's beta at the first k-site >= 5, which are
    adjacent visited sites, so the reference's mismatch-count matrix can be read back from stepping stones."""
    N = hap.shape[1]
    tau = theta / (1 - theta)
    for k in range(N):
        a = res["alpha"][k, 1].astype(np.float64)
        b = res["beta"][k, 0].astype(np.float64)
        s_b = res["site_end"][k, 0]
        assert res["site_begin"][k, 1] < 5 <= s_b
        for n in range(N):
            if n == k:
                continue
            mis_b = int(hap[s_b, k] == ord("1") and hap[s_b, n] == ord("0"))
            val = (np.log(a[n]) + np.log(b[n]) - np.log((1 - theta) / (N - 1))) / np.log(tau) + mis_b
            assert abs(val - KAT_D[k, n]) < 1e-3, (k, n, val)


def test_known_answer_5x5_mismatch_counts():
    hap, r, wb, theta = kat_inputs()
    res = oracle.paint_targets(hap, r, wb, theta, 0, 5)
    kat_check(res, hap, theta)


def test_rle_rule():
    v = np.array([1.0, 1.0005, 1.0009, 1.002, 0.0, 0.0, 5.0, 5.004, 5.006], np.float32)
    vals, lens = oracle.rle_encode(v)
    # runs are measured against the run HEAD; zeros never merge (collapsed_matrix.hpp:239-250)
    assert vals.tolist() == [np.float32(1.0), np.float32(1.002), 0.0, 0.0, np.float32(5.0), np.float32(5.006)]
    assert lens.tolist() == [3, 1, 1, 1, 2, 1]


def test_oracle_vs_reference_binary_fresh(tmp_path, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/Relate not built")
    d = str(tmp_path / "o")
    synth.make_chunk_dir(d, 120, 1500, seed=21, n_windows=7)
    oracle.run_reference(["--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", "0.001,1"], cwd=str(tmp_path))
    ref_dir = str(tmp_path / "ref_paint")
    os.rename(os.path.join(d, "chunk_0", "paint"), ref_dir)
    os.rmdir(os.path.join(d, "chunk_0"))
    oracle.paint_chunk(d, 0, "0.001,1")
    for w in range(7):
        assert filecmp.cmp(os.path.join(ref_dir, f"relate_{w}.bin"), os.path.join(d, "chunk_0", "paint", f"relate_{w}.bin"),
                           shallow=False)


def test_chunk_writer_matches_reference_makechunks(tmp_path, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/Relate not built")
    N, L, mem = 60, 900, 0.0006
    hap, bp = synth.block_kingman(N, L, 3)
    hp, sp = chunkio.write_haps_sample(str(tmp_path / "d"), hap, bp)
    chunkio.write_uniform_map(str(tmp_path / "map.txt"), bp)
    oracle.run_reference(["--mode", "MakeChunks", "--haps", hp, "--sample", sp, "--map", "map.txt", "-o", "ref",
                          "--memory", str(mem)], cwd=str(tmp_path))
    wb = chunkio.window_boundaries(hap, mem)
    assert len(wb) > 3
    chunkio.write_chunk(str(tmp_path / "mine"), hap, bp, chunkio.uniform_map_rpos(bp), wb)
    for f in ["parameters_c0.bin"] + [f"chunk_0.{e}" for e in ("hap", "bp", "dist", "r", "rpos", "state")]:
        assert filecmp.cmp(str(tmp_path / "ref" / f), str(tmp_path / "mine" / f), shallow=False), f


def test_paint_file_reader_roundtrip(tmp_path):
    d = unpack_golden("synth_n96", str(tmp_path))
    ch = chunkio.read_chunk(d, 0)
    recs = chunkio.read_paint_file(os.path.join(GOLDEN, "synth_n96", "paint_ref", "relate_2.bin"), ch.N)
    assert len(recs) == ch.N
    res = oracle.paint_targets(ch.hap, ch.r, ch.wb, float(np.float32(0.001)), 0, ch.N)
    for k, (a, b, ra, rb) in enumerate(recs):
        assert (a, b) == (ch.wb[2], ch.wb[3] - 1)
        assert ra.site == res["site_begin"][k, 2] and rb.site == res["site_end"][k, 2]
        assert ra.site <= a and rb.site >= b  # reader contract, anc_builder.cpp:67-69
        assert ra.logscale == res["ls_alpha"][k, 2] and rb.logscale == res["ls_beta"][k, 2]
        # decoded values are the run heads: within the codec's 1e-3 of the pre-RLE vector
        pre = res["alpha"][k, 2]
        dec = ra.expand()
        nz = pre != 0
        assert np.all(np.abs(dec[nz] - pre[nz]) <= 1.001e-3 * np.minimum(dec[nz], pre[nz]) + 1e-45)
