"""CPU tests of the C-ABI boundary: the library loads, exports what include/relate_paint.h declares, has no
CPU fallback, and its host-side pieces (RLE encoder, chunk loader errors, sharding logic) behave."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, unpack_golden
from oracle import oracle
from relate_b200 import capi


def header_symbols():
    src = open(os.path.join(ROOT, "include", "relate_paint.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/relate_paint.h but not exported"
    assert set(names) == set(capi.SYMBOLS), "ctypes binding and header disagree"


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "relate_paint.h")).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def gpu_count():
    return capi.lib().rp_device_count()


def test_no_cpu_fallback_without_a_device(tmp_path):
    if gpu_count() > 0:
        pytest.skip("a GPU is present")
    hap = np.full((10, 4), ord("0"), np.uint8)
    with pytest.raises(capi.PaintError) as e:
        capi.DeviceChunk.from_arrays(hap, np.full(10, 1e-3), np.array([0, 10], np.int32))
    assert e.value.code in (-4, -3)
    d = unpack_golden("synth_n96", str(tmp_path))
    with pytest.raises(capi.PaintError):
        capi.paint_chunk(d, 0, "0.001,1")
    # the CLI fails loudly too
    exe = os.path.join(ROOT, "relate_b200", "bin", "relate")
    p = subprocess.run([exe, "--mode", "Paint", "--chunk_index", "0", "-o", "out"], cwd=str(tmp_path), capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU path" in p.stderr


def test_argument_validation_needs_no_device():
    hap = np.full((10, 4), ord("0"), np.uint8)
    r = np.full(10, 1e-3)
    for wb in ([0, 9], [1, 10], [0, 5, 5, 10]):
        with pytest.raises(capi.PaintError) as e:
            capi.DeviceChunk.from_arrays(hap, r, np.array(wb, np.int32))
        assert e.value.code == -1
    with pytest.raises(capi.PaintError) as e:
        capi.DeviceChunk.from_arrays(hap, r, np.array([0, 10], np.int32), theta=1.5)
    assert e.value.code == -1


def test_chunk_load_reports_missing_files(tmp_path):
    with pytest.raises(capi.PaintError) as e:
        capi.DeviceChunk.load(str(tmp_path), 0)
    assert e.value.code == -2 and "parameters_c0.bin" in str(e.value)
    d = unpack_golden("synth_n96", str(tmp_path))
    os.remove(os.path.join(d, "chunk_0.rpos"))  # the reference loader opens all six files
    with pytest.raises(capi.PaintError) as e:
        capi.DeviceChunk.load(d, 0)
    assert e.value.code == -2 and "rpos" in str(e.value)


def test_host_rle_encoder_matches_oracle():
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 1000):
        for _ in range(20):
            base = rng.random(max(1, n // 7) + 1).astype(np.float32)
            v = np.repeat(base, 7)[:n] * (1 + rng.normal(0, 4e-4, n)).astype(np.float32)
            v[rng.random(n) < 0.05] = 0
            a, b = capi.rle_encode(v), oracle.rle_encode(v)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            assert a[1].sum() == n


@pytest.mark.parametrize("N,L", [(1, 3), (8, 70), (31, 9), (32, 5), (33, 5), (100, 40), (1000, 7), (1056, 3)])
def test_host_bit_packer(N, L):
    """The stage driver's reader threads pack the genotype rows they read (SSE2 compare + movemask, a scalar tail, zero
    padding): against numpy's packbits, for full words, partial last words and padded rows."""
    rng = np.random.default_rng(N * 131 + L)
    hap = np.where(rng.random((L, N)) < 0.4, ord("1"), ord("0")).astype(np.uint8)
    wps = (((N + 31) // 32 + 3) // 4) * 4
    G = np.full((L, wps), 0xDEADBEEF, np.uint32)
    capi.check(capi.lib().rp_debug_pack_host(N, L, hap.ctypes.data, G.ctypes.data, wps))
    exp = np.zeros((L, wps * 32), bool)
    exp[:, :N] = hap == ord("1")
    expG = np.packbits(exp.reshape(L, -1, 32), axis=-1, bitorder="little").view(np.uint32).reshape(L, -1)
    assert np.array_equal(G, expG)


def test_cli_flag_surface(tmp_path):
    exe = os.path.join(ROOT, "relate_b200", "bin", "relate")
    p = subprocess.run([exe, "--mode", "Paint"], capture_output=True, text=True)
    assert p.returncode == 0 and "Needed: chunk_index, output." in p.stdout  # Relate.cpp:66-76
    p = subprocess.run([exe, "--mode", "Paint", "-o", "a/b", "--chunk_index", "0"], capture_output=True, text=True)
    assert p.returncode == 1 and "Output needs to be in working directory." in p.stderr  # Relate.cpp:50-58
    p = subprocess.run([exe, "--mode", "Paint", "--no_such_flag", "1"], capture_output=True, text=True)
    assert p.returncode != 0
    # every reference flag is accepted (scripts pass a fixed set)
    flags = ["--haps", "h", "--sample", "s", "--map", "m", "-m", "1.25e-8", "-N", "30000", "--memory", "5", "--seed", "1",
             "--dist", "d", "--annot", "a", "--sample_ages", "x", "--coal", "c", "--fb", "1", "--no_consistency",
             "--transversion", "--first_section", "0", "--last_section", "1", "-i", "in", "--painting", "0.001,1",
             "--resident", "--gpu_topology", "--fp64"]
    p = subprocess.run([exe, "--mode", "Paint"] + flags, capture_output=True, text=True)
    assert p.returncode == 0 and "Needed: chunk_index, output." in p.stdout


def test_cli_delegates_other_modes_to_reference(tmp_path, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/Relate not built")
    exe = os.path.join(ROOT, "relate_b200", "bin", "relate")
    env = dict(os.environ, RELATE_REFERENCE_BIN=oracle.REF_RELATE)
    p = subprocess.run([exe, "--mode", "MakeChunks"], capture_output=True, text=True, env=env)
    assert "Needed: haps, sample, map, output." in p.stdout  # the reference's own usage text


