"""BASELINE.json configs[4] ("multi-chunk chromosome, 20 chunks, ~1M SNPs, chunks distributed over 8 B200 vs the
reference's Paint on all host cores"), Paint stage, with bounded disk.

    python scripts/config5_run.py [--N 2000] [--chunks 20] [--snps-per-chunk 50000] [--workdir /dev/shm] [--keep-log path]

What runs:
  1. a synthetic chromosome (block-Kingman, uniform map) is written as SHAPEIT haps/sample/map text;
  2. this repo's MakeChunks (rp_make_chunks_ex with the hapbits sidecar) cuts it into chunks (--memory chosen so that a
     chunk takes `snps-per-chunk` new SNPs; every chunk but the first also holds the 20 000-SNP overlap);
  3. Paint: rp_paint_chunks distributes whole chunks over all visible GPUs, in waves of one chunk per GPU; after each wave the
     paint files are deleted (what InferBranchLengths does in the pipeline, InferBranchLengths.cpp:60-75; the cluster
     scripts bound the number of live paintings the same way, RelateSlurm.sh:314), so at most n_gpus paintings are on disk;
  4. CPU baseline: min(cores, chunks) concurrent single-threaded `Relate --mode Paint` processes of the unmodified
     reference (what RelateParallel.sh's disabled `parallelize $chunks` would do, :416-425,548-549) on sample chunks
     of the same N (first `--cpu-snps` SNPs of the chromosome), rate extrapolated to the full chunks and labelled so.

N defaults to 2000, not BASELINE's 5000: at N = 5000 one chunk's paint files are ~100 GB (500 windows x 2 x N^2 x 4 B), so
live paintings need hundreds of GB of (RAM) disk: `--N 5000 --memory 1.74 --max-live 4 --cpu-snps 600` is the at-size run
(--memory 1.74 makes the 500-window cap close a chunk after ~50 000 new SNPs: 20 chunks).  Stated in the output.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from relate_b200 import capi, chunkio, synth  # noqa: E402
from oracle import oracle  # noqa: E402

PAINTING = "0.001,1"


def write_inputs(prefix, N, L, block=100000):
    """haps/sample/map text for a block-Kingman chromosome of L SNPs, generated and written block by block."""
    hp, sp, mp = prefix + ".haps", prefix + ".sample", prefix + ".map"
    with open(sp, "w") as f:
        f.write("ID_1 ID_2 missing\n0 0 0\n")
        for i in range(N // 2):
            f.write(f"s{i} s{i} 0\n")
    pos0, last = 0, 0
    with open(hp, "wb", buffering=1 << 24) as f:
        for b0 in range(0, L, block):
            n = min(block, L - b0)
            hap, bp = synth.block_kingman(N, n, seed=1000 + b0 // block)
            bp = bp.astype(np.int64) + pos0
            pos0 = int(bp[-1]) + 7
            body = np.full((n, 2 * N), ord(" "), np.uint8)
            body[:, 1::2] = hap
            for s in range(n):
                f.write(b"1 snp%d %d A T" % (b0 + s, bp[s]))
                f.write(body[s].tobytes())
                f.write(b"\n")
            last = int(bp[-1])
    top = last + 2
    with open(mp, "w") as f:
        f.write("pos COMBINED_rate Genetic_Map\n")
        f.write("0 1.0 0\n")
        f.write(f"{top} 1.0 {top * 1e-6!r}\n")
    return hp, sp, mp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=2000)
    ap.add_argument("--chunks", type=int, default=20)
    ap.add_argument("--snps-per-chunk", type=int, default=50000)
    ap.add_argument("--workdir", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ap.add_argument("--cpu-snps", type=int, default=3000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--memory", type=float, default=0.0, help="--memory for MakeChunks (default: what makes a chunk take snps-per-chunk new SNPs)")
    ap.add_argument("--max-live", type=int, default=0, help="paintings alive at a time (default: one per GPU)")
    ap.add_argument("--keep-log", default=None)
    args = ap.parse_args()
    N, C, per = args.N, args.chunks, args.snps_per_chunk
    L = C * per
    # chunk capacity = (memory*1e9/4 - (2N^2+3N)) / N new SNPs (data.cpp:129-139)
    memory = args.memory if args.memory > 0 else (per * N + 2.0 * N * N + 3.0 * N) * 4.0 / 1e9 * 1.000001
    ndev = capi.lib().rp_device_count()
    live = min(ndev, args.max_live) if args.max_live > 0 else ndev
    tmp = tempfile.mkdtemp(prefix="relate_config5_", dir=args.workdir)
    out = {"config": f"config 5{' (shrunk: BASELINE.json names N=5000)' if N < 5000 else ''}: N={N} x L={L} SNPs, --memory {memory:.4f} "
                     f"(aiming at {C} chunks of {per} new SNPs + the 20000-SNP overlap), --painting {PAINTING}; {ndev} GPU(s), "
                     f"at most {live} paintings alive",
           "workdir": tmp}
    try:
        t0 = time.perf_counter()
        hp, sp, mp = write_inputs(os.path.join(tmp, "in"), N, L)
        out["s_generate_text"] = time.perf_counter() - t0
        out["haps_bytes"] = os.path.getsize(hp)
        t0 = time.perf_counter()
        n_chunks, warn = capi.make_chunks(hp, sp, mp, os.path.join(tmp, "o"), memory_gb=memory, hapbits=True)
        out["s_makechunks"] = time.perf_counter() - t0
        out["n_chunks"] = n_chunks
        os.remove(hp)
        odir = os.path.join(tmp, "o")
        cells, windows = 0.0, []
        for c in range(n_chunks):
            hdr = np.fromfile(os.path.join(odir, f"parameters_c{c}.bin"), "<i4", 3)
            cells += float(hdr[0]) * hdr[0] * hdr[1]
            windows.append(int(hdr[2]) - 1)
        out["painted_cells"] = cells
        out["windows_per_chunk"] = [min(windows), max(windows)]
        # ---- Paint, waves of one chunk per GPU, paint files deleted after each wave ----
        waves, paint_bytes, t_paint = [], 0, 0.0
        agg = {k: 0.0 for k in ("ms_paint", "ms_prep", "ms_rle", "ms_d2h", "ms_write", "ms_load")}
        for c0 in range(0, n_chunks, live):
            c1 = min(n_chunks, c0 + live) - 1
            t0 = time.perf_counter()
            st = capi.paint_chunks(odir, c0, c1, PAINTING, devices=list(range(ndev)))
            dt = time.perf_counter() - t0
            t_paint += dt
            nb = 0
            for c in range(c0, c1 + 1):
                pd = os.path.join(odir, f"chunk_{c}", "paint")
                nb += sum(os.path.getsize(os.path.join(pd, f)) for f in os.listdir(pd))
                shutil.rmtree(os.path.join(odir, f"chunk_{c}"))
            paint_bytes += nb
            for k in agg:
                agg[k] = max(agg[k], st[k]) if k != "ms_paint" else agg[k] + st[k]
            waves.append({"chunks": [c0, c1], "seconds": dt, "paint_file_bytes": nb, "ms_paint_busiest_gpu": st["ms_paint"],
                          "ms_write": st["ms_write"]})
            print(f"wave chunks {c0}-{c1}: {dt:.2f} s, {nb / 1e9:.1f} GB of paint files, kernels {st['ms_paint']:.0f} ms on the busiest GPU", flush=True)
        assert not any(f.endswith(".hapbits") for f in os.listdir(odir)), "sidecars must be consumed"
        out.update({"s_paint_all_chunks": t_paint, "cells_per_s": cells / t_paint, "paint_file_bytes_total": paint_bytes,
                    "max_live_paintings": live, "ms_paint_kernels_sum_of_busiest_gpu_per_wave": agg["ms_paint"], "waves": waves})
        # ---- CPU baseline: concurrent reference Paint processes on sample chunks ----
        if not args.no_cpu and oracle.have_reference():
            cores = os.cpu_count() or 1
            procs = max(1, min(cores, n_chunks))
            Ls = args.cpu_snps
            base = os.path.join(tmp, "cpu_base")
            synth.make_chunk_dir(base, N, Ls, seed=1000, memory_gb=max(memory, 0.05))
            dirs = []
            for i in range(procs):
                d = os.path.join(tmp, f"cpu{i}", "o")
                shutil.copytree(base, d)
                dirs.append(d)
            t0 = time.perf_counter()
            ps = [subprocess.Popen([oracle.REF_RELATE, "--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", PAINTING],
                                   cwd=os.path.dirname(d), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in dirs]
            rcs = [p.wait() for p in ps]
            dt = time.perf_counter() - t0
            assert all(rc == 0 for rc in rcs), rcs
            rate = procs * float(N) * N * Ls / dt
            out["cpu_baseline"] = {"kind": "reference", "cores": procs, "host_cores": cores, "seconds": dt,
                                   "sample": f"{procs} concurrent single-threaded Relate --mode Paint processes, each on a chunk of N={N} x {Ls} SNPs",
                                   "cells_per_s": rate, "extrapolated_s_for_all_chunks": cells / rate,
                                   "note": "rate measured on the sample, extrapolated to the chromosome's chunks (a full run is hours of CPU)"}
            out["speedup_vs_cpu_extrapolated"] = (cells / rate) / t_paint
        print(json.dumps(out), flush=True)
        if args.keep_log:
            with open(args.keep_log, "w") as f:
                json.dump(out, f, indent=1)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
