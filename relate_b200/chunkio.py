"""On-disk formats at the Paint boundary (little-endian, x86-64).

Writers produce exactly what the reference's ``Data::MakeChunks`` writes for a
single-chunk data set (``/root/reference/include/src/data.cpp:117-518``), so that the
reference binary, the oracle and the CUDA path all read the same bytes; readers decode
``chunk_<c>/paint/relate_<w>.bin`` (``include/src/collapsed_matrix.hpp:228-296``,
``include/src/fast_painting.cpp:589-601``).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

LOWER_BOUND_R = 1e-10  # data.cpp:4


def uniform_map_rpos(bp: np.ndarray, cm_per_mb: float = 1.0) -> np.ndarray:
    """rpos[L+1] in Morgans for a two-point uniform map (0,0)-(bp_max+2, ...), evaluated with the
    reference's interpolation expression (data.cpp:456-463).  bp_pos[L] = bp_pos[L-1]+1 (data.cpp:351)."""
    bp = np.asarray(bp, dtype=np.int64)
    bpx = np.concatenate([bp, bp[-1:] + 1])
    map_bp0, map_bp1 = 0, int(bpx[-1]) + 1
    g0, g1 = 0.0, map_bp1 * 1e-6 * cm_per_mb
    return ((bpx - map_bp0) / float(map_bp1 - map_bp0) * (g1 - g0) + g0) * 1e-2


def write_uniform_map(path: str, bp: np.ndarray, cm_per_mb: float = 1.0) -> None:
    """Genetic-map text file matching :func:`uniform_map_rpos` (parsed by data.cpp:605-640)."""
    top = int(bp[-1]) + 2
    with open(path, "w") as f:
        f.write("pos COMBINED_rate Genetic_Map\n")
        f.write(f"0 {cm_per_mb} 0\n")
        f.write(f"{top} {cm_per_mb} {top * 1e-6 * cm_per_mb!r}\n")


def r_from_rpos(rpos: np.ndarray) -> np.ndarray:
    """data.cpp:466-477: r[s] = max(rpos[s+1]-rpos[s], 1e-10) * 2500."""
    d = rpos[1:] - rpos[:-1]
    d = np.where(d < LOWER_BOUND_R, LOWER_BOUND_R, d)
    return d * 2500


def window_boundaries(hap: np.ndarray, memory_gb: float) -> np.ndarray:
    """Window plan of Data::MakeChunks for a data set that fits one chunk (data.cpp:129-231).

    hap: [L, N] uint8 of '0'/'1' chars.  Returns wb[W+1] with wb[0]=0, wb[W]=L.
    """
    L, N = hap.shape
    mem = np.float32(memory_gb)
    min_memory_size = float(mem) * 1e9 / 4.0 - (2.0 * N * N + 3.0 * N)
    if min_memory_size <= 0:
        raise ValueError("Need larger memory allowance")
    max_chunk = min(L + 1, int(min_memory_size / N))
    if float(mem) >= 100:
        max_chunk = 2500000
    if L > max_chunk:
        raise ValueError("data set needs more than one chunk; use the reference MakeChunks")
    nder = (hap == ord("1")).sum(axis=1).astype(np.float64) * (N + 1)
    wb = [0]
    acc, in_win = 0.0, 0
    for s in range(L):
        if len(wb) >= 500:
            raise ValueError("more than 500 windows; use the reference MakeChunks")
        acc += nder[s]
        if acc >= min_memory_size and in_win > 10:
            in_win, acc = 0, 0.0
            wb.append(s)
        in_win += 1
    wb.append(L)
    return np.asarray(wb, dtype=np.int32)


def write_chunk(out_dir: str, hap: np.ndarray, bp: np.ndarray, rpos: np.ndarray, wb: np.ndarray,
                chunk: int = 0) -> None:
    """Write parameters_c<c>.bin and chunk_<c>.{hap,bp,dist,r,rpos,state} (+ parameters.bin for chunk 0)."""
    L, N = hap.shape
    os.makedirs(out_dir, exist_ok=True)
    base = os.path.join(out_dir, f"chunk_{chunk}")
    with open(os.path.join(out_dir, f"parameters_c{chunk}.bin"), "wb") as f:
        f.write(struct.pack("<iii", N, L, len(wb)))
        f.write(np.asarray(wb, dtype="<i4").tobytes())
    with open(base + ".hap", "wb") as f:
        f.write(struct.pack("<QQ", L, N))
        f.write(np.ascontiguousarray(hap, dtype=np.uint8).tobytes())
    bp = np.asarray(bp, dtype="<i4")
    dist = np.empty(L, dtype="<i4")
    dist[:-1] = bp[1:] - bp[:-1]
    dist[-1] = 1
    r = r_from_rpos(np.asarray(rpos, dtype=np.float64))
    for ext, hdr, arr in (("bp", L, bp), ("dist", L, dist), ("rpos", L + 1, np.asarray(rpos, "<f8")),
                          ("r", L, r.astype("<f8"))):
        with open(f"{base}.{ext}", "wb") as f:
            f.write(struct.pack("<I", hdr))
            f.write(arr.tobytes())
    with open(base + ".state", "wb") as f:
        f.write(struct.pack("<i", L))
        f.write(np.ones(L, dtype="<i4").tobytes())
    if chunk == 0:
        with open(os.path.join(out_dir, "parameters.bin"), "wb") as f:
            f.write(struct.pack("<iiid", N, L, 1, 0.0))
            f.write(struct.pack("<ii", 0, L))


def write_haps_sample(prefix: str, hap: np.ndarray, bp: np.ndarray) -> tuple[str, str]:
    """SHAPEIT haps/sample text (all diploid, ID_1==ID_2 so N = 2*rows; data.hpp:137-143)."""
    L, N = hap.shape
    assert N % 2 == 0
    haps_path, sample_path = prefix + ".haps", prefix + ".sample"
    with open(sample_path, "w") as f:
        f.write("ID_1 ID_2 missing\n0 0 0\n")
        for i in range(N // 2):
            f.write(f"s{i} s{i} 0\n")
    sep = np.full((L, N), ord(" "), dtype=np.uint8)
    body = np.empty((L, 2 * N), dtype=np.uint8)
    body[:, 0::2] = sep
    body[:, 1::2] = hap
    with open(haps_path, "wb") as f:
        for s in range(L):
            f.write(f"1 snp{s} {int(bp[s])} A T".encode())
            f.write(body[s].tobytes())
            f.write(b"\n")
    return haps_path, sample_path


@dataclass
class Chunk:
    N: int
    L: int
    wb: np.ndarray      # int32 [W+1]
    hap: np.ndarray     # uint8 [L, N] chars
    r: np.ndarray       # float64 [L]
    rpos: np.ndarray | None = None  # float64 [L+1]

    @property
    def W(self) -> int:
        return len(self.wb) - 1


def read_chunk(out_dir: str, chunk: int = 0) -> Chunk:
    with open(os.path.join(out_dir, f"parameters_c{chunk}.bin"), "rb") as f:
        N, L, nb = struct.unpack("<iii", f.read(12))
        wb = np.frombuffer(f.read(4 * nb), dtype="<i4").copy()
    base = os.path.join(out_dir, f"chunk_{chunk}")
    with open(base + ".hap", "rb") as f:
        uL, uN = struct.unpack("<QQ", f.read(16))
        assert (uL, uN) == (L, N)
        hap = np.frombuffer(f.read(L * N), dtype=np.uint8).reshape(L, N).copy()
    with open(base + ".r", "rb") as f:
        (n,) = struct.unpack("<I", f.read(4))
        assert n == L
        r = np.frombuffer(f.read(8 * L), dtype="<f8").copy()
    with open(base + ".rpos", "rb") as f:
        (n,) = struct.unpack("<I", f.read(4))
        assert n == L + 1
        rpos = np.frombuffer(f.read(8 * (L + 1)), dtype="<f8").copy()
    return Chunk(N, L, wb, hap, r, rpos)


@dataclass
class PaintRecord:
    site: int
    logscale: np.float32
    vals: np.ndarray    # float32 [K]
    lens: np.ndarray    # int32 [K]

    def expand(self) -> np.ndarray:
        return np.repeat(self.vals, self.lens)


def read_paint_file(path: str, N: int):
    """Decode one relate_<w>.bin -> list over targets of (start, end, alpha: PaintRecord, beta: PaintRecord)."""
    out = []
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def rec(off):
        one, sub = struct.unpack_from("<QQ", buf, off)
        assert one == 1 and sub == N, (one, sub)
        site, ls, k = struct.unpack_from("<ifi", buf, off + 16)
        off += 28
        vals = np.frombuffer(buf, dtype="<f4", count=k, offset=off).copy()
        off += 4 * k
        lens = np.frombuffer(buf, dtype="<i4", count=k, offset=off).copy()
        off += 4 * k
        assert int(lens.sum()) == N
        return PaintRecord(site, np.float32(ls), vals, lens), off

    while off < len(buf):
        a, b = struct.unpack_from("<ii", buf, off)
        off += 8
        ra, off = rec(off)
        rb, off = rec(off)
        out.append((a, b, ra, rb))
    return out
