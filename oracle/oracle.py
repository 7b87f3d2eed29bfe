"""ctypes wrapper of oracle/liboracle_paint.so — TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never from relate_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_paint.so")
REF_RELATE = os.path.join(HERE, "_ref", "Relate")
REF_DLENS = os.path.join(HERE, "_ref", "dlens")
REF_RELATE_GPU = os.path.join(HERE, "_ref", "Relate_gpu")  # the reference with GetMatrix bound to the GPU (bt_gpu_shim.cpp)
_lib = None


def build(ref: bool = True) -> None:
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build(ref=False)
        l = C.CDLL(LIB)
        l.ro_fast_log.restype = C.c_float
        l.ro_fast_log.argtypes = [C.c_float]
        l.ro_rle_encode.restype = C.c_int
        l.ro_rle_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        l.ro_paint_target.restype = C.c_int
        l.ro_paint_target.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                      C.c_int] + [C.c_void_p] * 6
        l.ro_count_sites.restype = C.c_long
        l.ro_count_sites.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.ro_paint_chunk.restype = C.c_int
        l.ro_paint_chunk.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        _lib = l
    return _lib


def fast_log(x: np.ndarray) -> np.ndarray:
    l = lib()
    x = np.asarray(x, dtype=np.float32)
    return np.array([l.ro_fast_log(float(v)) for v in x.ravel()], dtype=np.float32).reshape(x.shape)


def rle_encode(v: np.ndarray):
    v = np.ascontiguousarray(v, dtype=np.float32)
    vals = np.empty(len(v), np.float32)
    lens = np.empty(len(v), np.int32)
    k = lib().ro_rle_encode(v.ctypes.data, len(v), vals.ctypes.data, lens.ctypes.data)
    return vals[:k].copy(), lens[:k].copy()


def paint_targets(hap: np.ndarray, r: np.ndarray, wb: np.ndarray, theta: float, k_begin: int, k_end: int):
    """fp64 restatement of PaintSteppingStones for targets [k_begin,k_end) -> dict of pre-RLE arrays."""
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    r = np.ascontiguousarray(r, dtype=np.float64)
    wb = np.ascontiguousarray(wb, dtype=np.int32)
    L, N = hap.shape
    W = len(wb) - 1
    T = k_end - k_begin
    out = dict(alpha=np.empty((T, W, N), np.float32), beta=np.empty((T, W, N), np.float32),
               ls_alpha=np.empty((T, W), np.float32), ls_beta=np.empty((T, W), np.float32),
               site_begin=np.empty((T, W), np.int32), site_end=np.empty((T, W), np.int32))
    l = lib()
    for i, k in enumerate(range(k_begin, k_end)):
        rc = l.ro_paint_target(hap.ctypes.data, N, L, r.ctypes.data, theta, wb.ctypes.data, W, k,
                               out["alpha"][i].ctypes.data, out["beta"][i].ctypes.data,
                               out["ls_alpha"][i].ctypes.data, out["ls_beta"][i].ctypes.data,
                               out["site_begin"][i].ctypes.data, out["site_end"][i].ctypes.data)
        if rc:
            raise RuntimeError(f"ro_paint_target({k}) failed: {rc}")
    return out


def count_sites(hap: np.ndarray, k: int) -> int:
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    L, N = hap.shape
    return int(lib().ro_count_sites(hap.ctypes.data, N, L, k))


def paint_chunk(out_dir: str, chunk: int, painting: str | None = None, k_begin: int = 0, k_end: int = -1):
    st = (C.c_double * 4)()
    rc = lib().ro_paint_chunk(out_dir.encode(), chunk, painting.encode() if painting is not None else None,
                              k_begin, k_end, st)
    if rc:
        raise RuntimeError(f"ro_paint_chunk failed: {rc}")
    return dict(N=int(st[0]), L=int(st[1]), W=int(st[2]), sites=int(st[3]))


def have_reference() -> bool:
    return os.access(REF_RELATE, os.X_OK)


def run_reference(args, cwd, check=True):
    """Run the unmodified reference binary (oracle/_ref/Relate) with the given CLI args."""
    return subprocess.run([REF_RELATE] + list(args), cwd=cwd, capture_output=True, text=True, check=check)


def window_distances(out_dir: str, chunk: int, section: int, stride: int, painting: str | None, out_path: str) -> None:
    """Oracle restatement of RePaintSection + GetMatrix over one window; same output format as oracle/_ref/dlens."""
    l = lib()
    l.ro_window_distances.restype = C.c_int
    l.ro_window_distances.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    rc = l.ro_window_distances(out_dir.encode(), chunk, section, stride, (painting or "-").encode(), out_path.encode())
    if rc:
        raise RuntimeError(f"ro_window_distances failed: {rc}")


def read_distances(path: str) -> dict:
    """-> {snp: float32 [N,N]} from a dlens / window_distances file (plain or .gz)."""
    import gzip
    import struct
    buf = (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")).read()
    N, cnt = struct.unpack_from("<ii", buf, 0)
    off, out = 8, {}
    for _ in range(cnt):
        (snp,) = struct.unpack_from("<i", buf, off)
        out[snp] = np.frombuffer(buf, "<f4", N * N, off + 4).reshape(N, N)
        off += 4 + 4 * N * N
    return out
