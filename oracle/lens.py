"""d_ij lens for ONE target row — TEST INFRASTRUCTURE ONLY (see oracle.py).

oracle/_ref/dlens drives the unmodified reference's GetTopologyWithRepaint + GetMatrix over a whole window, which
needs every target's posterior (the `--memory` budget: tens of GB at N = 5 000).  Row n of the distance matrix
depends only on target n's own stepping stones, so at BASELINE.json's larger configs the parity checks look through
the same lens one row at a time, with the C restatements that tests/test_repaint_cpu.py pins byte-for-byte to dlens:
ro_repaint_section (FastPainting::RePaintSection, src/fast_painting.cpp:620-1092) and ro_matrix_row
(DistanceMeasure::GetMatrix, src/anc_builder.cpp:108-207).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle


def _bind():
    l = oracle.lib()
    l.ro_repaint_section.restype = C.c_int
    l.ro_repaint_section.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    l.ro_matrix_row.restype = None
    l.ro_matrix_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p]
    return l


def collapse(v: np.ndarray) -> np.ndarray:
    """What a reader of the paint files sees of a stepping stone: every element replaced by the head of its run
    (CollapsedMatrix<float>::DumpToFile then ReadFromFile, src/collapsed_matrix.hpp:228-296)."""
    vals, lens = oracle.rle_encode(np.ascontiguousarray(v, np.float32))
    return np.repeat(vals, lens).astype(np.float32)


def dij_rows(hap, r, rpos, wb, theta, w, n, snps, alpha_begin, beta_end, bS, eS, ls_alpha, ls_beta) -> np.ndarray:
    """Rows d[n][:] for the given SNPs of window w, from target n's DECODED stepping stones of that window.
    -> float32 [len(snps), N]."""
    l = _bind()
    hap = np.ascontiguousarray(hap, np.uint8)
    r = np.ascontiguousarray(r, np.float64)
    rpos = np.ascontiguousarray(rpos, np.float64)
    L, N = hap.shape
    W = len(wb) - 1
    start = int(wb[w])
    end = int(wb[w + 1]) - 1 if w < W - 1 else L - 1
    ab = np.ascontiguousarray(alpha_begin, np.float32)
    be = np.ascontiguousarray(beta_end, np.float32)
    assert bS <= start and eS >= end, (bS, start, eS, end)
    top = np.empty((eS - bS + 2, N), np.float32)
    ls = np.empty(eS - bS + 2, np.float32)
    l.ro_repaint_section(hap.ctypes.data, N, L, r.ctypes.data, float(theta), ab.ctypes.data, be.ctypes.data, int(bS),
                         int(eS), float(ls_alpha), float(ls_beta), int(n), top.ctypes.data, ls.ctypes.data)
    col = hap[:, n] == ord("1")
    out = np.empty((len(snps), N), np.float32)
    for i, snp in enumerate(snps):
        snp = int(snp)
        assert start <= snp <= end
        v = int(col[max(start, 1):snp + 1].sum())          # n-derived sites in [start, snp], SNP 0 not counted
        prev = np.flatnonzero(col[:snp + 1])
        rp_prev = float(rpos[prev[-1]]) if len(prev) else float(rpos[0])
        rp_next = C.c_double(rp_prev)                      # forces the search for the next derived site from snp
        l.ro_matrix_row(hap.ctypes.data, N, L, rpos.ctypes.data, snp, int(n), top.ctypes.data, ls.ctypes.data, v,
                        rp_prev, C.byref(rp_next), out[i].ctypes.data)
    return out


def paint_file_index(path: str, N: int):
    """Byte offsets of the N targets' blocks in a relate_<w>.bin (each block: int,int, alpha record, beta record)."""
    import struct
    offs = []
    with open(path, "rb") as f:
        off = 0
        for _ in range(N):
            offs.append(off)
            f.seek(off + 8 + 24)
            (ka,) = struct.unpack("<i", f.read(4))
            off2 = off + 8 + 28 + 8 * ka
            f.seek(off2 + 24)
            (kb,) = struct.unpack("<i", f.read(4))
            off = off2 + 28 + 8 * kb
    return offs


def read_target_records(path: str, N: int, n: int, index=None):
    """-> (alpha [N], site_begin, ls_alpha, beta [N], site_end, ls_beta) of target n, decoded from relate_<w>.bin."""
    import struct
    index = index or paint_file_index(path, N)
    with open(path, "rb") as f:
        f.seek(index[n] + 8)

        def rec():
            one, sub, site, ls, k = struct.unpack("<QQifi", f.read(28))
            assert one == 1 and sub == N
            vals = np.frombuffer(f.read(4 * k), "<f4")
            lens = np.frombuffer(f.read(4 * k), "<i4")
            return np.repeat(vals, lens).astype(np.float32), site, np.float32(ls)
        a, sa, la = rec()
        b, sb, lb = rec()
    return a, sa, la, b, sb, lb
