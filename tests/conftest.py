import gzip
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CHUNK_EXT = ["hap", "bp", "dist", "r", "rpos", "state"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def unpack_golden(name: str, dst_root: str, out: str = "out") -> str:
    """Materialise a golden chunk (stored gzipped as chunk 0) under <dst_root>/<out>/; returns that dir."""
    src = os.path.join(GOLDEN, name)
    d = os.path.join(dst_root, out)
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(src, "parameters_c0.bin"), d)
    for e in CHUNK_EXT:
        with gzip.open(os.path.join(src, f"chunk_0.{e}.gz"), "rb") as g, open(os.path.join(d, f"chunk_0.{e}"), "wb") as f:
            f.write(g.read())
    return d


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b|/|b| over entries where either is non-zero."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    m = (a != 0) | (b != 0)
    if not m.any():
        return 0.0
    return float((np.abs(a - b)[m] / np.maximum(np.abs(b[m]), 1e-300)).max())


def make_case(N, L, W, seed):
    from relate_b200 import chunkio, synth
    hap, bp = synth.block_kingman(N, L, seed)
    rpos = chunkio.uniform_map_rpos(bp)
    r = chunkio.r_from_rpos(rpos)
    wb = np.linspace(0, L, W + 1).astype(np.int32)
    wb[0], wb[-1] = 0, L
    return hap, r, wb


@pytest.fixture(scope="session")
def have_ref():
    from oracle import oracle
    return oracle.have_reference()
