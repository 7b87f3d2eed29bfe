"""rp_make_chunks (host loader) against the unmodified reference's `Relate --mode MakeChunks`:
every output file byte for byte, on the bundled example (gzip input, multi-chunk with --memory small enough to
split... the example is 130k SNPs x 8 haplotypes, so the 500-window cap and the 20000-SNP overlap are exercised),
on synthetic text haps with a multi-row genetic map, a dist file and --transversion, plus the golden md5s (which
need no reference at run time)."""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import oracle
from relate_b200 import capi, chunkio, synth

EXE = os.path.join(ROOT, "relate_b200", "bin", "relate")


def write_haps(d, hap, bp, alleles=None, gz=False, diploid=True):
    """hap [L,N] of '0'/'1' bytes -> SHAPEIT haps/sample text files; returns (haps, sample) paths."""
    L, N = hap.shape
    rng = np.random.default_rng(7)
    if alleles is None:
        alleles = [("ACGT"[i % 4], "ACGT"[(i + 1 + int(rng.integers(3))) % 4]) for i in range(L)]
    lines = []
    for s in range(L):
        a, b = alleles[s]
        lines.append(f"1 snp{s} {int(bp[s])} {a} {b} " + " ".join(chr(c) for c in hap[s]) + "\n")
    hp = os.path.join(d, "x.haps" + (".gz" if gz else ""))
    (gzip.open(hp, "wt") if gz else open(hp, "w")).writelines(lines)
    sp = os.path.join(d, "x.sample")
    with open(sp, "w") as f:
        f.write("ID_1 ID_2 missing\n0 0 0\n")
        if diploid:
            for i in range(N // 2):
                f.write(f"s{i} s{i} 0\n")
        else:  # haploid rows have differing ids (data.hpp:137-143)
            for i in range(N):
                f.write(f"s{i} NA 0\n")
    return hp, sp


def write_map(d, bp, rows=40, seed=3):
    rng = np.random.default_rng(seed)
    xs = np.unique(np.concatenate([[max(0, int(bp[0]) - 50)], np.sort(rng.integers(int(bp[0]), int(bp[-1]), rows)), [int(bp[-1]) + 10]]))
    g = np.cumsum(rng.random(len(xs)) * 0.05)
    p = os.path.join(d, "map.txt")
    with open(p, "w") as f:
        f.write("pos COMBINED_rate Genetic_Map\n")
        for x, y in zip(xs, g):
            f.write(f"{int(x)} {rng.random():.4f} {y:.9f}\n")
    return p


def dir_md5(d):
    return {n: hashlib.md5(open(os.path.join(d, n), "rb").read()).hexdigest() for n in sorted(os.listdir(d))}


def run_ref(cwd, args):
    subprocess.run([oracle.REF_RELATE, "--mode", "MakeChunks", "-o", "ref"] + args, cwd=cwd, check=True, capture_output=True)
    return dir_md5(os.path.join(cwd, "ref"))


CASES = [  # N, L, seed, memory, extra
    (20, 900, 1, 0.0001, {}),
    (64, 2500, 2, 0.001, {"transversion": True}),
    (30, 1500, 3, 5.0, {"dist": True}),
    (12, 300, 4, 0.00005, {"haploid": True, "gz": True}),
]


@pytest.mark.parametrize("N,L,seed,mem,extra", CASES)
def test_synthetic_matches_reference(tmp_path, have_ref, N, L, seed, mem, extra):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path)
    hap, bp = synth.block_kingman(N, L, seed)
    hp, sp = write_haps(d, hap, bp, gz=extra.get("gz", False), diploid=not extra.get("haploid", False))
    mp = write_map(d, bp)
    args = ["--haps", hp, "--sample", sp, "--map", mp, "--memory", repr(mem)]
    dist = None
    if extra.get("dist"):
        dist = os.path.join(d, "x.dist")
        with open(dist, "w") as f:
            f.write("#pos dist\n")
            for s in range(L):
                f.write(f"{int(bp[s])} {int(bp[s + 1] - bp[s]) if s + 1 < L else 7}\n")
        args += ["--dist", dist]
    if extra.get("transversion"):
        args += ["--transversion"]
    ref = run_ref(d, args)
    n, warn = capi.make_chunks(hp, sp, mp, os.path.join(d, "mine"), dist=dist, transversion=bool(extra.get("transversion")), memory_gb=mem)
    mine = dir_md5(os.path.join(d, "mine"))
    assert mine == ref
    assert n == sum(1 for k in ref if k.startswith("parameters_c"))
    assert "hard disc" in warn


def test_bundled_example_multichunk_matches_reference(tmp_path, have_ref):
    if not have_ref or not os.path.exists("/root/reference/example/data/example.haps.gz"):
        pytest.skip("needs the reference checkout")
    d = str(tmp_path)
    ex = "/root/reference/example/data/"
    with gzip.open(ex + "example.haps.gz", "rt") as f:
        last = int(f.readlines()[-1].split(" ", 3)[2])
    mp = os.path.join(d, "map.txt")
    with open(mp, "w") as f:
        f.write("pos COMBINED_rate Genetic_Map\n0 1.0 0\n%d 1.0 %r\n" % (last + 2, (last + 2) * 1e-6))
    args = ["--haps", ex + "example.haps.gz", "--sample", ex + "example.sample.gz", "--map", mp, "--memory", "0.001"]
    ref = run_ref(d, args)
    # through the drop-in CLI this time
    subprocess.run([EXE, "--mode", "MakeChunks", "-o", "mine"] + args, cwd=d, check=True, capture_output=True)
    assert dir_md5(os.path.join(d, "mine")) == ref
    assert sum(1 for k in ref if k.startswith("parameters_c")) == 5


def test_golden_md5_without_reference(tmp_path):
    """Fixture written by tests/golden/make_golden.py from the reference's MakeChunks output."""
    g = json.load(open(os.path.join(GOLDEN, "makechunks_synth.json")))
    d = str(tmp_path)
    hap, bp = synth.block_kingman(g["N"], g["L"], g["seed"])
    hp, sp = write_haps(d, hap, bp)
    mp = write_map(d, bp)
    capi.make_chunks(hp, sp, mp, os.path.join(d, "mine"), memory_gb=g["memory"])
    assert dir_md5(os.path.join(d, "mine")) == g["md5"]


def test_errors_do_not_exit(tmp_path):
    d = str(tmp_path)
    hap, bp = synth.block_kingman(10, 50, 1)
    hp, sp = write_haps(d, hap, bp)
    mp = write_map(d, bp)
    os.makedirs(os.path.join(d, "exists"))
    with pytest.raises(capi.PaintError) as e:  # pipeline/MakeChunks.cpp:38-43
        capi.make_chunks(hp, sp, mp, os.path.join(d, "exists"))
    assert "already exists" in str(e.value)
    with pytest.raises(capi.PaintError):
        capi.make_chunks(hp + ".nope", sp, mp, os.path.join(d, "o1"))
    with pytest.raises(capi.PaintError) as e:  # data.cpp:127-130
        capi.make_chunks(hp, sp, mp, os.path.join(d, "o2"), memory_gb=1e-9)
    assert "larger memory" in str(e.value)
    # unsorted positions (data.cpp:393-397)
    bp2 = bp.copy()
    bp2[10] = bp2[9]
    hp2, sp2 = write_haps(os.path.join(d), hap, bp2)
    with pytest.raises(capi.PaintError) as e:
        capi.make_chunks(hp2, sp2, mp, os.path.join(d, "o3"))
    assert "not sorted" in str(e.value)


def test_hapbits_sidecar_holds_the_packed_rows_and_changes_nothing_else(tmp_path):
    """RP_MC_HAPBITS: chunk_<c>.hapbits = the rows of chunk_<c>.hap in the painter's bit layout (bit n&31 of word n>>5,
    rows padded to 16 bytes); every other file is what MakeChunks writes without the option."""
    import struct
    d = str(tmp_path)
    N, L = 70, 1200
    hap, bp = synth.block_kingman(N, L, 9)
    hp, sp = write_haps(d, hap, bp)
    mp = write_map(d, bp)
    capi.make_chunks(hp, sp, mp, os.path.join(d, "plain"), memory_gb=0.001)
    capi.make_chunks(hp, sp, mp, os.path.join(d, "bits"), memory_gb=0.001, hapbits=True)
    a, b = dir_md5(os.path.join(d, "plain")), dir_md5(os.path.join(d, "bits"))
    side = {k: v for k, v in b.items() if k.endswith(".hapbits")}
    assert side and {k: v for k, v in b.items() if k not in side} == a
    for name in side:
        c = int(name.split("_")[1].split(".")[0])
        ch = chunkio.read_chunk(os.path.join(d, "bits"), c)
        buf = open(os.path.join(d, "bits", name), "rb").read()
        magic, n, l, wps, _ = struct.unpack_from("<8siiii", buf, 0)
        assert magic == b"RPHBITS1" and (n, l) == (ch.N, ch.L) and wps == ((ch.N + 31) // 32 + 3) // 4 * 4
        rows = np.frombuffer(buf, "<u4", l * wps, 24).reshape(l, wps)
        want = np.zeros((l, wps * 32), np.uint8)
        want[:, :ch.N] = ch.hap == ord("1")
        assert np.array_equal(rows, np.packbits(want, axis=1, bitorder="little").view("<u4"))
