"""ctypes binding of ``include/relate_paint.h`` (``librelate_paint.so``).

This is the only way Python code in this repo reaches the painting path: there is no
Python/numpy/torch implementation of it and no CPU fallback.  If the library is missing or
no CUDA device is present, calls raise :class:`PaintError`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RELATE_PAINT_LIB") or os.path.join(_HERE, "librelate_paint.so")

RP_FP64 = 1

ERRORS = {-1: "EINVAL", -2: "EIO", -3: "ECUDA", -4: "ENODEVICE", -5: "ENOMEM", -6: "EUNSUPPORTED"}


class PaintError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"relate_paint: {ERRORS.get(code, code)}: {msg}")
        self.code = code


class RpInfo(C.Structure):
    _fields_ = [("N", C.c_int), ("L", C.c_int), ("W", C.c_int), ("device", C.c_int),
                ("words_per_snp", C.c_int), ("hbm_bytes", C.c_longlong)]


class RpStats(C.Structure):
    _fields_ = [("ms_h2d", C.c_double), ("ms_prep", C.c_double), ("ms_paint", C.c_double),
                ("ms_d2h", C.c_double), ("ms_encode", C.c_double), ("ms_total", C.c_double),
                ("sites", C.c_longlong), ("cells", C.c_longlong), ("h2d_bytes", C.c_longlong),
                ("d2h_bytes", C.c_longlong), ("launches", C.c_int), ("n_targets", C.c_int),
                ("team_threads", C.c_int), ("words_per_thread", C.c_int), ("ctas", C.c_int),
                ("reserved", C.c_int), ("ms_load", C.c_double), ("ms_rle", C.c_double),
                ("ms_write", C.c_double)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class RpMinMatchStats(C.Structure):
    _fields_ = [("ms_kernel", C.c_float), ("draws", C.c_longlong), ("first_fallback_step", C.c_int),
                ("fallback_steps", C.c_int), ("general_steps", C.c_int), ("medium_steps", C.c_int), ("launches", C.c_int)]


class RpTune(C.Structure):
    _fields_ = [("words_per_thread", C.c_int), ("ctas_per_sm", C.c_int), ("reserved", C.c_int * 6)]


# every symbol include/relate_paint.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "rp_last_error": (C.c_char_p, []),
    "rp_device_count": (C.c_int, []),
    "rp_version": (C.c_char_p, []),
    "rp_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "rp_host_free": (None, [_P]),
    "rp_chunk_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_double, C.c_uint, C.POINTER(_P)]),
    "rp_chunk_load": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_uint, C.POINTER(_P)]),
    "rp_chunk_info": (C.c_int, [_P, C.POINTER(RpInfo)]),
    "rp_chunk_free": (None, [_P]),
    "rp_chunk_set_tune": (C.c_int, [_P, C.POINTER(RpTune)]),
    "rp_chunk_set_stream": (C.c_int, [_P, _P]),
    "rp_peak_fp32": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rp_paint_targets": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.POINTER(RpStats)]),
    "rp_paint_targets_device": (C.c_int, [_P, C.c_int, C.c_int] + [C.POINTER(_P)] * 6 + [C.POINTER(RpStats)]),
    "rp_paint_records": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(RpStats)]),
    "rp_records_copy": (C.c_int, [_P, C.c_longlong, C.c_longlong, _P, C.POINTER(RpStats)]),
    "rp_paint_from_host": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_double, C.c_uint,
                                     C.POINTER(RpTune), C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.POINTER(RpStats)]),
    "rp_paint_chunk": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, _P, C.c_int, C.c_uint, C.POINTER(RpStats)]),
    "rp_paint_chunks": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_char_p, _P, C.c_int, C.c_uint, C.POINTER(RpStats)]),
    "rp_stage_device_stats": (C.c_int, [C.c_int, C.POINTER(RpStats)]),
    "rp_release_cache": (None, []),
    "rp_window_open": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.POINTER(_P), C.POINTER(RpStats)]),
    "rp_window_open_resident": (C.c_int, [_P, C.c_int, _P, C.POINTER(_P), C.POINTER(RpStats)]),
    "rp_window_open_files": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(RpStats)]),
    "rp_window_distance": (C.c_int, [_P, C.c_int, _P]),
    "rp_window_rows": (C.c_longlong, [_P]),
    "rp_window_close": (None, [_P]),
    "rp_make_chunks": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_float,
                                 C.POINTER(C.c_int), C.c_char_p, C.c_size_t]),
    "rp_make_chunks_ex": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_float, C.c_uint,
                                    C.POINTER(C.c_int), C.c_char_p, C.c_size_t]),
    "rp_minmatch_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.POINTER(_P)]),
    "rp_minmatch_create_thresholds": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(_P)]),
    "rp_minmatch_destroy": (None, [_P]),
    "rp_minmatch_reset": (C.c_int, [_P]),
    "rp_minmatch_quickbuild": (C.c_int, [_P, _P, _P, _P, C.POINTER(RpMinMatchStats)]),
    "rp_minmatch_quickbuild_device": (C.c_int, [_P, _P, _P, _P, C.POINTER(RpMinMatchStats)]),
    "rp_rle_encode": (C.c_int, [_P, C.c_int, _P, _P]),
    "rp_fast_log_device": (C.c_int, [C.c_int, _P, _P, C.c_int]),
    "rp_debug_pack_host": (C.c_int, [C.c_int, C.c_int, _P, _P, C.c_int]),
    "rp_debug_pack": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_int), _P, C.POINTER(C.c_int)]),
}

_lib = None


def lib():
    """Load the CUDA library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PaintError(-4, f"{LIB_PATH} not built (run __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise PaintError(rc, lib().rp_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class SteppingStones:
    """Pre-RLE output of painting targets [k_begin, k_end)."""
    k_begin: int
    alpha: np.ndarray       # float32 [T, W, N]
    beta: np.ndarray        # float32 [T, W, N]
    ls_alpha: np.ndarray    # float32 [T, W]
    ls_beta: np.ndarray     # float32 [T, W]
    site_begin: np.ndarray  # int32 [T, W]
    site_end: np.ndarray    # int32 [T, W]
    stats: dict


class DeviceChunk:
    """A chunk resident in one GPU's HBM (``rp_chunk``)."""

    def __init__(self, handle, keep=()):
        self._h = handle
        self._keep = keep
        info = RpInfo()
        check(lib().rp_chunk_info(self._h, C.byref(info)))
        self.N, self.L, self.W, self.device = info.N, info.L, info.W, info.device
        self.words_per_snp, self.hbm_bytes = info.words_per_snp, info.hbm_bytes

    @classmethod
    def from_arrays(cls, hap: np.ndarray, r: np.ndarray, wb: np.ndarray, theta: float = 0.001,
                    device: int = 0, fp64: bool = False) -> "DeviceChunk":
        hap = np.ascontiguousarray(hap, dtype=np.uint8)
        r = np.ascontiguousarray(r, dtype=np.float64)
        wb = np.ascontiguousarray(wb, dtype=np.int32)
        L, N = hap.shape
        h = C.c_void_p()
        check(lib().rp_chunk_create(device, N, L, _ptr(hap), _ptr(r), _ptr(wb), len(wb), theta,
                                    RP_FP64 if fp64 else 0, C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, out_dir: str, chunk_index: int, painting: str | None = None, device: int = 0,
             fp64: bool = False) -> "DeviceChunk":
        h = C.c_void_p()
        check(lib().rp_chunk_load(device, out_dir.encode(), chunk_index,
                                  painting.encode() if painting is not None else None,
                                  RP_FP64 if fp64 else 0, C.byref(h)))
        return cls(h)

    def set_tune(self, words_per_thread: int = 0, ctas_per_sm: int = 0, cluster: int = 0, segments: int = 0) -> None:
        """cluster > 1 forces teams of that many CTAs (thread-block cluster) where a single CTA would do (tests);
        segments: 0 = automatic, 1 = whole chains as jobs, n = every chain cut into n parked segments."""
        t = RpTune(words_per_thread, ctas_per_sm)
        t.reserved[1] = segments
        t.reserved[3] = cluster
        check(lib().rp_chunk_set_tune(self._h, C.byref(t)))

    def set_stream(self, cuda_stream: int | None) -> None:
        """Issue the chunk's work on an external CUDA stream (e.g. torch.cuda.Stream().cuda_stream)."""
        check(lib().rp_chunk_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def paint_targets(self, k_begin: int = 0, k_end: int | None = None, vectors: bool = True) -> SteppingStones:
        k_end = self.N if k_end is None else k_end
        T = k_end - k_begin
        alpha = np.empty((T, self.W, self.N), np.float32) if vectors else None
        beta = np.empty((T, self.W, self.N), np.float32) if vectors else None
        lsa = np.empty((T, self.W), np.float32)
        lsb = np.empty((T, self.W), np.float32)
        sb = np.empty((T, self.W), np.int32)
        se = np.empty((T, self.W), np.int32)
        st = RpStats()
        check(lib().rp_paint_targets(self._h, k_begin, k_end, _ptr(alpha), _ptr(beta), _ptr(lsa), _ptr(lsb),
                                     _ptr(sb), _ptr(se), C.byref(st)))
        return SteppingStones(k_begin, alpha, beta, lsa, lsb, sb, se, st.as_dict())

    def paint_targets_device(self, k_begin: int = 0, k_end: int | None = None) -> dict:
        """Paint with results left in HBM; returns the stats (device pointers stay inside the library)."""
        k_end = self.N if k_end is None else k_end
        ptrs = [C.c_void_p() for _ in range(6)]
        st = RpStats()
        check(lib().rp_paint_targets_device(self._h, k_begin, k_end, *[C.byref(p) for p in ptrs], C.byref(st)))
        d = st.as_dict()
        d["dev_ptrs"] = [p.value for p in ptrs]
        return d

    def paint_records(self, k_begin: int = 0, k_end: int | None = None):
        """Paint + device record encoder: -> (list of W ``bytes`` file images for these targets, stats)."""
        k_end = self.N if k_end is None else k_end
        off = np.zeros(self.W + 1, np.int64)
        st = RpStats()
        check(lib().rp_paint_records(self._h, k_begin, k_end, _ptr(off), C.byref(st)))
        img = np.empty(int(off[-1]), np.uint8)
        check(lib().rp_records_copy(self._h, 0, int(off[-1]), _ptr(img), C.byref(st)))
        return [img[off[w]:off[w + 1]].tobytes() for w in range(self.W)], st.as_dict()

    def close(self) -> None:
        if self._h:
            lib().rp_chunk_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Window:
    """Posterior rows of one window resident in HBM (``rp_window``): the GPU side of RePaintSection + GetMatrix."""

    def __init__(self, chunk: DeviceChunk, handle, stats):
        self._c, self._h, self.stats = chunk, handle, stats
        self.rows = lib().rp_window_rows(handle)

    @classmethod
    def open_files(cls, chunk: DeviceChunk, out_dir: str, chunk_index: int, w: int) -> "Window":
        h, st = C.c_void_p(), RpStats()
        check(lib().rp_window_open_files(chunk._h, out_dir.encode(), chunk_index, w, C.byref(h), C.byref(st)))
        return cls(chunk, h, st.as_dict())

    @classmethod
    def open_resident(cls, chunk: DeviceChunk, w: int, rpos) -> "Window":
        """From the stepping stones of the last ``paint_targets_device(0, N)`` still in HBM (no paint files)."""
        rpos = np.ascontiguousarray(rpos, np.float64)
        h, st = C.c_void_p(), RpStats()
        check(lib().rp_window_open_resident(chunk._h, w, _ptr(rpos), C.byref(h), C.byref(st)))
        return cls(chunk, h, st.as_dict())

    @classmethod
    def open(cls, chunk: DeviceChunk, w: int, alpha, beta, ls_alpha, ls_beta, rpos) -> "Window":
        arrs = [np.ascontiguousarray(alpha, np.float32), np.ascontiguousarray(beta, np.float32),
                np.ascontiguousarray(ls_alpha, np.float32), np.ascontiguousarray(ls_beta, np.float32),
                np.ascontiguousarray(rpos, np.float64)]
        h, st = C.c_void_p(), RpStats()
        check(lib().rp_window_open(chunk._h, w, *[_ptr(a) for a in arrs], C.byref(h), C.byref(st)))
        return cls(chunk, h, st.as_dict())

    def distance(self, snp: int, out: np.ndarray | None = None) -> np.ndarray:
        """N x N distance matrix for one SNP.  ``out`` may be a preallocated float32 array; pinned memory
        (:func:`pinned_empty`) lets the device->host copy run at PCIe speed instead of through a staging buffer."""
        d = np.empty((self._c.N, self._c.N), np.float32) if out is None else out
        assert d.dtype == np.float32 and d.shape == (self._c.N, self._c.N) and d.flags.c_contiguous
        check(lib().rp_window_distance(self._h, snp, _ptr(d)))
        return d

    def close(self) -> None:
        if self._h:
            lib().rp_window_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class MinMatch:
    """One tree builder (``rp_minmatch_*``) = one reference ``MinMatch`` object: state survives from tree to tree."""

    def __init__(self, N: int, theta: float, device: int = 0):
        self.N = N
        self._h = _P()
        check(lib().rp_minmatch_create(device, N, theta, C.byref(self._h)))

    def quickbuild(self, d: np.ndarray, prior: np.ndarray | None = None):
        """-> (merges int32 [N-1, 2], stats dict); d / prior: float32 [N, N] on the host, not modified."""
        N = self.N
        d = np.ascontiguousarray(d, dtype=np.float32)
        assert d.shape == (N, N)
        if prior is not None:
            prior = np.ascontiguousarray(prior, dtype=np.float32)
            assert prior.shape == (N, N)
        merges = np.empty((N - 1, 2), np.int32)
        st = RpMinMatchStats()
        check(lib().rp_minmatch_quickbuild(self._h, _ptr(d), _ptr(prior), _ptr(merges), C.byref(st)))
        return merges, {n: getattr(st, n) for n, _ in st._fields_}

    def quickbuild_device(self, dev_d: int, dev_prior: int | None = None):
        """Same with the matrices already in the handle's device memory (raw device pointers, N x N float32 each)."""
        merges = np.empty((self.N - 1, 2), np.int32)
        st = RpMinMatchStats()
        check(lib().rp_minmatch_quickbuild_device(self._h, _P(dev_d), _P(dev_prior) if dev_prior else None, _ptr(merges), C.byref(st)))
        return merges, {n: getattr(st, n) for n, _ in st._fields_}

    def reset(self):
        """Back to the state of a freshly constructed MinMatch object."""
        check(lib().rp_minmatch_reset(self._h))

    def close(self):
        if self._h:
            lib().rp_minmatch_destroy(self._h)
            self._h = _P()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def paint_from_host(hap, r, wb, theta=0.001, device=0, fp64=False, k_begin=0, k_end=None, out=None,
                    words_per_thread=0, ctas_per_sm=0) -> SteppingStones:
    """chunk_create + paint + copy back + free in one C call (the end-to-end call bench.py times).

    ``out`` may carry preallocated (e.g. pinned) arrays: dict with alpha,beta,ls_alpha,ls_beta,site_begin,site_end.
    """
    L, N = hap.shape
    W = len(wb) - 1
    k_end = N if k_end is None else k_end
    T = k_end - k_begin
    if out is None:
        out = dict(alpha=np.empty((T, W, N), np.float32), beta=np.empty((T, W, N), np.float32),
                   ls_alpha=np.empty((T, W), np.float32), ls_beta=np.empty((T, W), np.float32),
                   site_begin=np.empty((T, W), np.int32), site_end=np.empty((T, W), np.int32))
    st = RpStats()
    tune = RpTune(words_per_thread, ctas_per_sm)
    check(lib().rp_paint_from_host(device, N, L, _ptr(hap), _ptr(r), _ptr(wb), len(wb), theta,
                                   RP_FP64 if fp64 else 0, C.byref(tune), k_begin, k_end, _ptr(out["alpha"]),
                                   _ptr(out["beta"]), _ptr(out["ls_alpha"]), _ptr(out["ls_beta"]),
                                   _ptr(out["site_begin"]), _ptr(out["site_end"]), C.byref(st)))
    return SteppingStones(k_begin, out["alpha"], out["beta"], out["ls_alpha"], out["ls_beta"],
                          out["site_begin"], out["site_end"], st.as_dict())


def paint_chunk(out_dir: str, chunk_index: int, painting: str | None = None, devices=None, fp64: bool = False) -> dict:
    """The Paint stage (``Relate --mode Paint``): writes ``<out_dir>/chunk_<c>/paint/relate_<w>.bin``."""
    st = RpStats()
    dev = None
    n = 0
    if devices is not None:
        dev = np.asarray(list(devices), dtype=np.int32)
        n = len(dev)
    check(lib().rp_paint_chunk(out_dir.encode(), chunk_index, painting.encode() if painting is not None else None,
                               _ptr(dev), n, RP_FP64 if fp64 else 0, C.byref(st)))
    return st.as_dict()


def stage_device_stats(n_devices: int) -> list:
    """Per-device statistics of the last :func:`paint_chunk` call of this process."""
    out = []
    for i in range(n_devices):
        st = RpStats()
        check(lib().rp_stage_device_stats(i, C.byref(st)))
        out.append(st.as_dict())
    return out


def paint_chunks(out_dir: str, first_chunk: int, last_chunk: int, painting: str | None = None, devices=None,
                 fp64: bool = False) -> dict:
    """Paint chunks first_chunk..last_chunk, whole chunks per device (``rp_paint_chunks``)."""
    st = RpStats()
    dev = None
    n = 0
    if devices is not None:
        dev = np.asarray(list(devices), dtype=np.int32)
        n = len(dev)
    check(lib().rp_paint_chunks(out_dir.encode(), first_chunk, last_chunk, painting.encode() if painting is not None else None,
                                _ptr(dev), n, RP_FP64 if fp64 else 0, C.byref(st)))
    return st.as_dict()


def make_chunks(haps: str, sample: str, gmap: str, out_dir: str, dist: str | None = None, transversion: bool = False,
                memory_gb: float = 5.0, hapbits: bool = False):
    """``Relate --mode MakeChunks`` (host-only): -> (number of chunks, the warnings the reference prints to stderr).
    hapbits: also write the bit-packed sidecar ``chunk_<c>.hapbits`` that :func:`paint_chunk` consumes."""
    n = C.c_int(0)
    buf = C.create_string_buffer(4096)
    check(lib().rp_make_chunks_ex(haps.encode(), sample.encode(), gmap.encode(), dist.encode() if dist else None,
                                  out_dir.encode(), int(transversion), memory_gb, 1 if hapbits else 0, C.byref(n), buf, len(buf)))
    return n.value, buf.value.decode()


def rle_encode(v: np.ndarray):
    v = np.ascontiguousarray(v, dtype=np.float32)
    vals = np.empty(len(v), np.float32)
    lens = np.empty(len(v), np.int32)
    k = lib().rp_rle_encode(_ptr(v), len(v), _ptr(vals), _ptr(lens))
    if k < 0:
        check(k)
    return vals[:k].copy(), lens[:k].copy()


def fast_log_device(x: np.ndarray, device: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    check(lib().rp_fast_log_device(device, _ptr(x), _ptr(out), x.size))
    return out


def peak_fp32(device: int = 0):
    """-> (lane-ops/s with the kernel's packed mix, lane-ops/s with scalar adds), measured on the device."""
    a, b = C.c_double(), C.c_double()
    check(lib().rp_peak_fp32(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def pinned_empty(shape, dtype):
    """numpy array backed by cudaMallocHost memory (freed when the array's owner is collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().rp_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED.append((p, buf))
    return arr


_PINNED: list = []
