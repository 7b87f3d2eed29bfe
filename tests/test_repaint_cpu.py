"""CPU tests pinning the oracle's restatement of RePaintSection + GetMatrix (oracle/repaint_oracle.c) to the reference:
bit-identical distance matrices vs fixtures written by oracle/_ref/dlens (the unmodified reference's GetMatrix) and
vs a fresh dlens run when the binary is present."""
import filecmp
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, unpack_golden
from oracle import oracle
from relate_b200 import synth


def test_oracle_distances_match_reference_golden(tmp_path):
    d = unpack_golden("synth_n96", str(tmp_path))
    os.makedirs(os.path.join(d, "chunk_0", "paint"))
    for w in range(5):
        shutil.copy(os.path.join(GOLDEN, "synth_n96", "paint_ref", f"relate_{w}.bin"), os.path.join(d, "chunk_0", "paint"))
    for sec in (1, 3):
        out = str(tmp_path / f"d_{sec}.bin")
        oracle.window_distances(d, 0, sec, 61, "0.001,1", out)
        want = gzip.open(os.path.join(GOLDEN, "synth_n96", "dlens_ref", f"d_{sec}.bin.gz"), "rb").read()
        assert open(out, "rb").read() == want


def test_oracle_distances_match_fresh_reference_run(tmp_path, have_ref):
    if not have_ref or not os.access(oracle.REF_DLENS, os.X_OK):
        pytest.skip("oracle/_ref not built")
    N, L, W = 90, 1600, 3
    synth.make_chunk_dir(str(tmp_path / "o"), N, L, seed=77, n_windows=W)
    oracle.run_reference(["--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", "0.002,1.5"], cwd=str(tmp_path))
    for sec in range(W):
        a, b = str(tmp_path / f"ref_{sec}.bin"), str(tmp_path / f"ora_{sec}.bin")
        subprocess.run([oracle.REF_DLENS, "o", "0", str(sec), "37", "0.002,1.5", a], cwd=str(tmp_path), check=True)
        oracle.window_distances(str(tmp_path / "o"), 0, sec, 37, "0.002,1.5", b)
        assert filecmp.cmp(a, b, shallow=False)
    dm = oracle.read_distances(b)
    assert len(dm) >= 3 and all(m.shape == (N, N) and np.all(np.diag(m) == 0) for m in dm.values())


def test_row_lens_reproduces_reference_golden_rows(tmp_path):
    """oracle/lens.py looks at one target row at a time (the at-size parity checks use it where the whole-window
    reference run would need the --memory budget in RAM): rows must be bit-identical to the matrices that the
    unmodified reference's GetMatrix wrote (dlens fixtures)."""
    from oracle import lens
    from relate_b200 import chunkio
    d = unpack_golden("synth_n96", str(tmp_path))
    ch = chunkio.read_chunk(d, 0)
    theta = float(np.float32(0.001))
    for sec in (1, 3):
        want = oracle.read_distances(os.path.join(GOLDEN, "synth_n96", "dlens_ref", f"d_{sec}.bin.gz"))
        pf = os.path.join(GOLDEN, "synth_n96", "paint_ref", f"relate_{sec}.bin")
        idx = lens.paint_file_index(pf, ch.N)
        snps = sorted(want)
        for n in (0, 17, ch.N - 1):
            a, sa, la, b, sb, lb = lens.read_target_records(pf, ch.N, n, idx)
            rows = lens.dij_rows(ch.hap, ch.r, ch.rpos, ch.wb, theta, sec, n, snps, a, b, sa, sb, la, lb)
            for i, snp in enumerate(snps):
                assert np.array_equal(rows[i], want[snp][n]), (sec, n, snp)
