"""Synthetic inputs of the shapes BASELINE.json names (block-Kingman haplotypes, uniform map).

The generator itself is C (``csrc/synth.c``, built by ``__graft_entry__.build()`` into
``_synth.so``); this module wraps it and writes MakeChunks-compatible chunk files.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import chunkio

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = ctypes.CDLL(path)
        _lib.synth_block_kingman.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                             ctypes.c_void_p, ctypes.c_void_p]
        _lib.synth_block_kingman.restype = ctypes.c_int
    return _lib


def block_kingman(N: int, L: int, seed: int, block: int = 50):
    """-> (hap uint8 [L,N] of '0'/'1' chars, bp int32 [L])."""
    hap = np.empty((L, N), dtype=np.uint8)
    bp = np.empty(L, dtype=np.int32)
    rc = _load().synth_block_kingman(N, L, block, seed, hap.ctypes.data, bp.ctypes.data)
    if rc:
        raise RuntimeError(f"synth_block_kingman failed: {rc}")
    return hap, bp


def make_chunk_dir(out_dir: str, N: int, L: int, seed: int, memory_gb: float = 5.0, n_windows: int | None = None):
    """Generate a single-chunk data set and write it as the reference's MakeChunks would.

    ``n_windows`` overrides the memory rule with W equal-length windows (for tests that want a
    particular W).  Returns (hap, bp, rpos, wb).
    """
    hap, bp = block_kingman(N, L, seed)
    rpos = chunkio.uniform_map_rpos(bp)
    if n_windows is None:
        wb = chunkio.window_boundaries(hap, memory_gb)
    else:
        wb = np.linspace(0, L, n_windows + 1).astype(np.int32)
        wb[0], wb[-1] = 0, L
    chunkio.write_chunk(out_dir, hap, bp, rpos, wb)
    return hap, bp, rpos, wb
