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


# ---- MinMatch (tree builder) : oracle restatement and the reference-linked tree lens --------------------------------
REF_QBLENS = os.path.join(HERE, "_ref", "qblens")


class MinMatchOracle:
    """oracle/minmatch_oracle.c: one MinMatch object (state survives from tree to tree as in the reference)."""

    def __init__(self, N: int, theta: float):
        l = lib()
        l.mmo_create.restype = C.c_void_p
        l.mmo_create.argtypes = [C.c_int, C.c_double]
        l.mmo_destroy.argtypes = [C.c_void_p]
        l.mmo_quickbuild.restype = C.c_int
        l.mmo_quickbuild.argtypes = [C.c_void_p] * 5
        self.N, self._l = N, l
        self._h = l.mmo_create(N, theta)

    def quickbuild(self, d: np.ndarray, prior: np.ndarray | None = None):
        """-> (merges int32 [N-1, 2], info dict).  d is not modified (a copy is)."""
        N = self.N
        dd = np.array(d, dtype=np.float32, order="C", copy=True).reshape(N, N)
        pp = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32).reshape(N, N)
        merges = np.empty((N - 1, 2), np.int32)
        info = (C.c_long * 2)()
        rc = self._l.mmo_quickbuild(self._h, dd.ctypes.data, None if pp is None else pp.ctypes.data, merges.ctypes.data, info)
        if rc:
            raise RuntimeError(f"mmo_quickbuild failed: {rc}")
        return merges, dict(draws=int(info[0]), first_sym_step=int(info[1]))

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.mmo_destroy(self._h)
            self._h = None


def write_tree_lens_input(path: str, N: int, theta: float, trees) -> None:
    """trees: list of (d, prior-or-None) -> the input format of oracle/_ref/qblens."""
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<idi", N, theta, len(trees)))
        for d, prior in trees:
            f.write(struct.pack("<i", 0 if prior is None else 1))
            f.write(np.ascontiguousarray(d, dtype="<f4").tobytes())
            if prior is not None:
                f.write(np.ascontiguousarray(prior, dtype="<f4").tobytes())


def reference_quickbuild(N: int, theta: float, trees, workdir: str, repeat: int = 1):
    """The reference's own MinMatch::QuickBuild (oracle/_ref/qblens) on a list of (d, prior) -> (list of merges, seconds)."""
    inp, out = os.path.join(workdir, "qb_in.bin"), os.path.join(workdir, "qb_out.bin")
    write_tree_lens_input(inp, N, theta, trees)
    r = subprocess.run([REF_QBLENS, inp, out, str(repeat)], check=True, capture_output=True, text=True)
    secs = float(r.stderr.strip().split("QuickBuild")[1].split()[0])
    m = np.fromfile(out, "<i4").reshape(len(trees), N - 1, 2)
    return [m[t] for t in range(len(trees))], secs


def prior_from_merges(merges: np.ndarray, N: int, val: float) -> np.ndarray:
    """The `dist` matrix AncesTreeBuilder::BuildTopology derives from the previous tree (src/anc_builder.cpp:581-606):
    dist[i][j] += val for every internal clade that contains i but not j."""
    members = [np.array([k]) for k in range(N)] + [None] * (N - 1)
    dist = np.zeros((N, N), np.float32)
    for t, (a, b) in enumerate(merges):
        mem = np.concatenate([members[a], members[b]])
        members[N + t] = mem
        add = np.full(N, np.float32(val), np.float32)
        add[mem] = 0
        dist[mem] += add[None, :]
    return dist
