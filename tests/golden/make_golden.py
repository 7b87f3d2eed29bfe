"""Generates the golden fixtures in this directory.  Run HERE (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

Fixtures (all produced by the UNMODIFIED reference binary oracle/_ref/Relate):
  example_c1/        chunk 1 of the reference's bundled example (example/data/example.{haps,sample}.gz, N=8),
                     cut by the reference's MakeChunks with --memory 0.001 and a substitute uniform 1 cM/Mb
                     map (the example's own map is a missing blob): parameters_c0.bin + chunk_0.* (gzipped),
                     the reference's paint files with --painting 0.001,1 (paint_ref/relate_<w>.bin) and the
                     md5 of the .anc/.mut files the reference's BuildTopology --seed 1 derives from them.
  synth_n96/         a 96-haplotype x 700-SNP synthetic chunk (this repo's generator, seed 5), 5 windows,
                     painted by the reference with --painting 0.001,1 and without the flag; dlens_ref/ holds the
                     distance matrices the reference's DistanceMeasure::GetMatrix derives from those paint files
                     (windows 1 and 3, every 61st SNP), written by oracle/_ref/dlens.
  makechunks_synth.json  md5 of every file the reference's MakeChunks writes for a 24 x 1200 synthetic text data set
                     (one chunk, several windows; inputs regenerated from the seed by tests/test_makechunks_cpu.py).
"""
import gzip
import hashlib
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REL = os.path.join(ROOT, "oracle", "_ref", "Relate")
EX = "/root/reference/example/data/"
CHUNK_EXT = ["hap", "bp", "dist", "r", "rpos", "state"]


def run(args, cwd):
    subprocess.run([REL] + args, cwd=cwd, check=True, capture_output=True)


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def pack_chunk(src_dir, c, dst):
    """store chunk c of src_dir as chunk 0 in dst (gzipped)"""
    os.makedirs(dst, exist_ok=True)
    shutil.copy(os.path.join(src_dir, f"parameters_c{c}.bin"), os.path.join(dst, "parameters_c0.bin"))
    for e in CHUNK_EXT:
        with open(os.path.join(src_dir, f"chunk_{c}.{e}"), "rb") as f, gzip.GzipFile(os.path.join(dst, f"chunk_0.{e}.gz"), "wb", 9, mtime=0) as g:
            g.write(f.read())


def main():
    import numpy as np
    from relate_b200 import chunkio, synth

    tmp = tempfile.mkdtemp()
    # ---- bundled example, chunk 1 -------------------------------------------------------
    pos = []
    with gzip.open(EX + "example.haps.gz", "rt") as f:
        for line in f:
            pos.append(int(line.split(" ", 3)[2]))
    top = pos[-1] + 2
    with open(os.path.join(tmp, "map.txt"), "w") as f:
        f.write("pos COMBINED_rate Genetic_Map\n0 1.0 0\n%d 1.0 %r\n" % (top, top * 1e-6))
    run(["--mode", "MakeChunks", "--haps", EX + "example.haps.gz", "--sample", EX + "example.sample.gz", "--map",
         "map.txt", "-o", "ex", "--memory", "0.001"], tmp)
    dst = os.path.join(HERE, "example_c1")
    shutil.rmtree(dst, ignore_errors=True)
    pack_chunk(os.path.join(tmp, "ex"), 1, dst)
    # work on a directory where it is chunk 0 (BuildTopology seeds depend on chunk_index)
    w1 = os.path.join(tmp, "w1")
    os.makedirs(os.path.join(w1, "ex"))
    shutil.copy(os.path.join(dst, "parameters_c0.bin"), os.path.join(w1, "ex"))
    for e in CHUNK_EXT:
        with gzip.open(os.path.join(dst, f"chunk_0.{e}.gz"), "rb") as g, open(os.path.join(w1, "ex", f"chunk_0.{e}"), "wb") as f:
            f.write(g.read())
    N, L, nb = struct.unpack("<iii", open(os.path.join(w1, "ex", "parameters_c0.bin"), "rb").read(12))
    W = nb - 1
    run(["--mode", "Paint", "--chunk_index", "0", "-o", "ex", "--painting", "0.001,1"], w1)
    os.makedirs(os.path.join(dst, "paint_ref"))
    for w in range(W):
        shutil.copy(os.path.join(w1, "ex", "chunk_0", "paint", f"relate_{w}.bin"), os.path.join(dst, "paint_ref"))
    run(["--mode", "BuildTopology", "--chunk_index", "0", "--first_section", "0", "--last_section", str(W - 1), "-o", "ex",
         "--painting", "0.001,1", "--seed", "1"], w1)
    sums = {}
    for w in range(W):
        for ext in ("anc", "mut"):
            p = os.path.join(w1, "ex", "chunk_0", f"ex_{w}.{ext}")
            sums[f"ex_{w}.{ext}"] = md5(p)
    json.dump({"N": N, "L": L, "W": W, "painting": "0.001,1", "seed": 1, "md5": sums},
              open(os.path.join(dst, "topology_md5.json"), "w"), indent=1)

    # ---- small synthetic ------------------------------------------------------------------
    dst = os.path.join(HERE, "synth_n96")
    shutil.rmtree(dst, ignore_errors=True)
    w2 = os.path.join(tmp, "w2")
    os.makedirs(w2)
    synth.make_chunk_dir(os.path.join(w2, "sy"), 96, 700, seed=5, n_windows=5)
    pack_chunk(os.path.join(w2, "sy"), 0, dst)
    for tag, extra in (("paint_ref", ["--painting", "0.001,1"]), ("paint_ref_noflag", [])):
        shutil.rmtree(os.path.join(w2, "sy", "chunk_0"), ignore_errors=True)
        run(["--mode", "Paint", "--chunk_index", "0", "-o", "sy"] + extra, w2)
        os.makedirs(os.path.join(dst, tag))
        for w in range(5):
            shutil.copy(os.path.join(w2, "sy", "chunk_0", "paint", f"relate_{w}.bin"), os.path.join(dst, tag))
    # d_ij lens outputs of the reference (oracle/_ref/dlens drives the reference's GetMatrix) on its own paint files
    shutil.rmtree(os.path.join(w2, "sy", "chunk_0"), ignore_errors=True)
    run(["--mode", "Paint", "--chunk_index", "0", "-o", "sy", "--painting", "0.001,1"], w2)
    os.makedirs(os.path.join(dst, "dlens_ref"))
    for sec in (1, 3):
        out = os.path.join(w2, f"d_{sec}.bin")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "dlens"), "sy", "0", str(sec), "61", "0.001,1", out], cwd=w2, check=True)
        with open(out, "rb") as f, gzip.GzipFile(os.path.join(dst, "dlens_ref", f"d_{sec}.bin.gz"), "wb", 9, mtime=0) as g:
            g.write(f.read())
    shutil.rmtree(tmp)
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(HERE) for f in fs)
    print("golden fixtures written,", total, "bytes")


def make_makechunks_fixture():
    """md5 of every file the reference's MakeChunks writes for a small synthetic text data set (the inputs are
    regenerated by tests/test_makechunks_cpu.py's writers from the seed)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_makechunks_cpu as t
    from relate_b200 import synth
    N, L, seed, mem = 24, 1200, 9, 0.0002
    tmp = tempfile.mkdtemp()
    hap, bp = synth.block_kingman(N, L, seed)
    hp, sp = t.write_haps(tmp, hap, bp)
    mp = t.write_map(tmp, bp)
    md5s = t.run_ref(tmp, ["--haps", hp, "--sample", sp, "--map", mp, "--memory", repr(mem)])
    json.dump({"N": N, "L": L, "seed": seed, "memory": mem, "md5": md5s}, open(os.path.join(HERE, "makechunks_synth.json"), "w"), indent=1)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "makechunks":  # only the MakeChunks fixture
        make_makechunks_fixture()
    else:
        main()
        make_makechunks_fixture()
