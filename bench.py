#!/usr/bin/env python
"""bench.py — painted cells/s of the B200-native `relate --mode Paint` hot path.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU Paint (oracle/_ref/Relate)

Workload (BASELINE.json configs[1]): synthetic block-Kingman haplotypes N=1000 x L=50000, one chunk, window
plan of `--memory 5`, `--painting 0.001,1`.  A step = one pass of the hot path over one chunk: per-target site
tables + the forward/backward kernel for all N targets.  With N GPUs every rank paints its own chunk of that
shape (independent chunks, no data-path collective): weak scaling, value = total painted cells / max-over-ranks time.

`value` is the device-resident rate (inputs already in HBM, stepping stones left in HBM); the quantity that is
comparable with the reference arm — chunk files in, paint files out — is `e2e`.  Both arms print the same `config`.

Keys beyond the base contract:
  roofline     dominant kernel (paint_kernel): algorithmic FP32 lane-ops 7*N*U per launch (SURVEY.md 8d;
               U = visited sites, counted exactly) / mean launch time (CUDA events on the launching stream),
               against the measured FP32 lane-op rate of this box (rp_peak_fp32: same instruction mix, no memory).
  e2e          the same metric through the whole reference-facing stage rp_paint_chunk (== `relate --mode Paint`):
               chunk files in, chunk_0/paint/relate_<w>.bin out; host->device and device->host copies, RLE encoding
               and file writes inside the timed region.
  cpu_baseline the unmodified reference binary on a bounded sample of the same workload, on this box's host.
  sharded      strong scaling of ONE chunk (BASELINE.json configs[2]: N=5000 x L=100000, --memory 50) whose targets
               are sharded over the N GPUs by rp_paint_chunk(devices=[0..N-1]), run from rank 0 while the other
               ranks wait on a CPU (gloo) barrier: stage wall time, per-device kernel time and roofline fraction of the
               multi-warp kernel, md5 of the paint files (must equal the 1-device files) and a d_ij check through
               the oracle's one-row lens.  `sharded_config4` (N=10000 x L=100000, --memory 100) is added at 8 GPUs
               (or RELATE_BENCH_CONFIG4=1) when the box has the disk and RAM for its 40 GB of paint files.
  window_repaint / e2e_resident   the consumer side (row f1): RePaintSection + GetMatrix of one window on the device, and the
               file-less path chunk files -> first distance matrix.
  tree_builder the consumer side (row f4), N=1 only: rp_minmatch_quickbuild per tree at N=1000 and N=5000 against the
               reference's own MinMatch::QuickBuild on one host core (oracle/_ref/qblens), merge lists checked identical,
               and the rate of 64 handles building N=1000 trees side by side (one tree per SM).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the tree_builder leg runs many one-CTA kernels side by side: streams share 8 hardware queues by default and kernels of streams
# that share one serialise (measured: 850 trees/s with 8, 2 900 with 32); must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_HAP, N_SNP, MEMORY_GB, PAINTING = 1000, 50000, 5.0, "0.001,1"
WORKLOAD = f"synthetic block-Kingman N={N_HAP} x L={N_SNP}, single chunk, --memory {MEMORY_GB:g}, --painting {PAINTING}"
METRIC, UNIT = "painted cells/s (N^2*L/s), relate --mode Paint", "cells/s"
# identical in both arms (the driver compares them)
CONFIG = {"workload": WORKLOAD,
          "l2": "GPU arm: flushed between timed iterations (256 MiB device write outside the event pair); reference arm: n/a",
          "timing": "GPU arm: torch.cuda.Event pairs on the stream the library launches on, max over ranks; "
                    "reference arm: host wall clock around the concurrent Paint processes"}
SHARDED = {"config3": dict(N=5000, L=100000, seed=2, memory=50.0),
           "config4": dict(N=10000, L=100000, seed=3, memory=100.0)}


def make_chunk(out_dir, seed, L=N_SNP):
    from relate_b200 import synth
    return synth.make_chunk_dir(out_dir, N_HAP, L, seed, memory_gb=MEMORY_GB)


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic():
    """DRAM bytes per paint_kernel launch from the latest committed `ncu --set full` summary (profiles/), or None."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_config2_paint_*_ncu_summary.csv"))):
        rd = wr = None
        for line in open(p):
            f = line.strip().split(",")
            if f[0] == "dram__bytes_read.sum":
                rd = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(f[1], 1)
            if f[0] == "dram__bytes_write.sum":
                wr = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(f[1], 1)
        if rd is not None and wr is not None:
            best = (rd + wr, os.path.basename(p))
    return best


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """The reference's own CPU Paint (oracle/_ref/Relate, built from /root/reference by oracle/Makefile) on this box's
    host cores: P concurrent single-threaded `Relate --mode Paint` processes (Paint has no threads, Paint.cpp:81-87;
    this is what RelateParallel.sh's disabled `parallelize $chunks` would do), each on its own bounded sample chunk
    of the workload's shape (N=1000, L_s SNPs).  Falls back to the oracle port if the binary did not travel."""
    if rank != 0:
        return
    from oracle import oracle
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    L_s = int(os.environ.get("RELATE_BENCH_REF_SNPS", "10000"))  # W > 1 windows, start-up cost amortised as in the workload
    use_ref = oracle.have_reference()
    tmp = tempfile.mkdtemp(prefix="relate_ref_")
    try:
        base = os.path.join(tmp, "base")
        make_chunk(base, seed=1, L=L_s)
        dirs = []
        for i in range(procs):
            d = os.path.join(tmp, f"p{i}", "o")
            shutil.copytree(base, d)
            dirs.append(d)

        def one_step():
            t0 = time.perf_counter()
            if use_ref:
                ps = [subprocess.Popen([oracle.REF_RELATE, "--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", PAINTING],
                                       cwd=os.path.dirname(d), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in dirs]
                rcs = [p.wait() for p in ps]
                assert all(rc == 0 for rc in rcs), rcs
            else:
                ts = [threading.Thread(target=oracle.paint_chunk, args=(d, 0, PAINTING)) for d in dirs]
                [t.start() for t in ts]
                [t.join() for t in ts]
            dt = time.perf_counter() - t0
            for d in dirs:
                shutil.rmtree(os.path.join(d, "chunk_0"), ignore_errors=True)
            return dt

        for _ in range(args.warmup):
            one_step()
        times = [one_step() for _ in range(args.steps)]
        ms = 1e3 * sum(times) / len(times)
        value = procs * N_HAP * N_HAP * L_s / (ms * 1e-3)
        kind = "reference" if use_ref else "port"
        sample = (f"{procs} concurrent single-threaded Paint processes, each N={N_HAP} x L={L_s} SNPs of the workload's "
                  f"generator (same --painting); {'oracle/_ref/Relate' if use_ref else 'oracle port'}")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": CONFIG, "sample": sample,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample,
                                 "host_cores": cores},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_sample():
    """Rank 0, N=1: the reference binary (1 core: Paint is single-threaded) on a bounded sample, ~10-20 s."""
    from oracle import oracle
    L_s = int(os.environ.get("RELATE_BENCH_CPU_SNPS", "10000"))
    tmp = tempfile.mkdtemp(prefix="relate_cpu_")
    try:
        d = os.path.join(tmp, "o")
        make_chunk(d, seed=1, L=L_s)
        t0 = time.perf_counter()
        if oracle.have_reference():
            subprocess.run([oracle.REF_RELATE, "--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", PAINTING],
                           cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            kind = "reference"
        else:
            oracle.paint_chunk(d, 0, PAINTING)
            kind = "port"
        dt = time.perf_counter() - t0
        return {"value": N_HAP * N_HAP * L_s / dt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
                "sample": f"one Paint of N={N_HAP} x L={L_s} SNPs (first {L_s} SNPs' worth of the workload's generator), "
                          f"1 core because the reference's Paint is single-threaded per chunk; host has {os.cpu_count()} cores"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def window_repaint_measure(chunk, rpos, W, peaks):
    """SURVEY.md 8 row f1 on the same chunk: RePaintSection for all N targets of one window (repaint_kernel, HBM-bound:
    the forward sweep writes every alpha row, the backward sweep reads it back and writes the posterior row) and
    GetMatrix for a few SNPs (distance_kernel), from the stepping stones still resident in HBM."""
    import numpy as np
    from relate_b200 import capi
    chunk.set_stream(None)
    chunk.paint_targets_device(0, N_HAP)
    w = W // 2
    times = []
    for rep in range(3):  # the first call also pays for the window's HBM allocations
        win = capi.Window.open_resident(chunk, w, rpos)
        times.append(win.stats["ms_paint"])
        rows = win.rows
        if rep < 2:
            win.close()
    ms = sorted(times)[1]
    lo = int(np.asarray(chunk_wb(chunk))[w])
    nd = 8
    t0 = time.perf_counter()
    for i in range(nd):
        win.distance(lo + 11 * i)
    ms_d = 1e3 * (time.perf_counter() - t0) / nd
    dpin = capi.pinned_empty((N_HAP, N_HAP), np.float32)
    win.distance(lo, out=dpin)
    t0 = time.perf_counter()
    for i in range(nd):
        win.distance(lo + 11 * i, out=dpin)
    ms_dp = 1e3 * (time.perf_counter() - t0) / nd
    win.close()
    ck = 4.0  # checkpoint spacing of the single-warp one-word teams (paint_api.cu)
    bytes_alg = 4.0 * N_HAP * rows + 2.0 * 4.0 * N_HAP * N_HAP            # compulsory: the posterior rows out, the stepping stones in
    bytes_design = (4.0 + 8.0 / ck) * N_HAP * rows + 2.0 * 4.0 * N_HAP * N_HAP  # + every ck-th alpha row written and read back
    peak = peaks.get("hbm_gbs") or 6500.0
    return {"window": w, "posterior_rows": int(rows), "repaint_kernel_ms": ms, "algorithmic_bytes": bytes_alg,
            "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9, "peak_gbs": peak, "frac": bytes_alg / (ms * 1e-3) / 1e9 / peak,
            "design_traffic_bytes": bytes_design, "design_traffic_gbs": bytes_design / (ms * 1e-3) / 1e9,
            "round1_design_traffic_bytes": 3.0 * 4.0 * N_HAP * rows + 2.0 * 4.0 * N_HAP * N_HAP,
            "bound": "hbm", "distance_call_ms": ms_d, "distance_call_pinned_ms": ms_dp,
            "distance_call": "rp_window_distance: distance_kernel + D2H of the N x N float matrix, host wall clock",
            "source": "stepping stones resident in HBM (rp_window_open_resident), no paint-file round trip"}


def chunk_wb(chunk):
    return chunk._wb


def resident_leg(out_dir, device, N, reps=3):
    """Chunk files -> first N x N distance matrix on the host without any paint file (the consumer-side product path of
    INTEGRATION.md section 4): rp_chunk_load + rp_paint_targets_device(all targets) + rp_window_open_resident(window 0)
    + rp_window_distance(first SNP of the window), host wall clock per repetition."""
    import numpy as np
    from relate_b200 import capi, chunkio
    rpos = chunkio.read_chunk(out_dir, 0).rpos
    d = capi.pinned_empty((N, N), np.float32)
    capi.lib().rp_release_cache()   # parked workspaces of earlier legs (their HBM is wanted for the window's posterior)
    runs, parts = [], {}
    for _ in range(1 + reps):
        t0 = time.perf_counter()
        with capi.DeviceChunk.load(out_dir, 0, PAINTING, device=device) as c:
            t1 = time.perf_counter()
            st = c.paint_targets_device(0, N)
            t2 = time.perf_counter()
            with capi.Window.open_resident(c, 0, rpos) as win:
                t3 = time.perf_counter()
                win.distance(0, out=d)
                t4 = time.perf_counter()
                runs.append(t4 - t0)        # (freeing the window's and the chunk's HBM afterwards is not part of it)
                rows, ms_rep = win.rows, win.stats["ms_paint"]
        parts = {"ms_load_h2d_pack": 1e3 * (t1 - t0), "ms_paint_call": 1e3 * (t2 - t1), "ms_paint_kernel": st["ms_paint"],
                 "ms_window_open": 1e3 * (t3 - t2), "ms_repaint_kernel": ms_rep, "ms_distance_call": 1e3 * (t4 - t3)}
    warm = sorted(runs[1:])
    assert np.isfinite(d).all() and float(np.abs(np.diag(d)).max()) == 0.0
    return {"call": "rp_chunk_load -> rp_paint_targets_device(0..N) -> rp_window_open_resident(0) -> rp_window_distance(snp 0): "
                    "chunk files in, first N x N distance matrix on the host, no paint files",
            "seconds": warm[(len(warm) - 1) // 2], "runs_ms": [round(1e3 * t, 2) for t in runs], "first_call_ms": 1e3 * runs[0],
            "window0_posterior_rows": int(rows), "breakdown_ms": parts}


def _md5_files(paths):
    """md5 over the concatenation order-independent digest list of the files (hashed in parallel threads)."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor

    def one(p):
        h = hashlib.md5()
        with open(p, "rb") as f:
            while True:
                b = f.read(1 << 24)
                if not b:
                    break
                h.update(b)
        return h.hexdigest()
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex:
        digs = list(ex.map(one, paths))
    return hashlib.md5("".join(digs).encode()).hexdigest(), sum(os.path.getsize(p) for p in paths)


def sharded_leg(name, devices, peaks, reps=3):
    """Strong scaling of one chunk: rp_paint_chunk(devices) from this process; see the module docstring."""
    import numpy as np
    from relate_b200 import capi, chunkio, synth
    from oracle import lens, oracle
    cfg = SHARDED[name]
    N, L = cfg["N"], cfg["L"]
    cells = float(N) * N * L
    nominal = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6
    # GBs of paint files per call: on a disk-backed file system the kernel's write-back of the dirty pages of earlier calls
    # throttles the later ones (measured: 190 -> 280 ms per config-3 call within one process); a RAM disk, when there is one
    # with room, times the stage itself.  RELATE_BENCH_TMP overrides.
    W_est = {"config3": 22, "config4": 40}[name]
    need_est = 2.2 * W_est * N * N * 4 + 2.0 * N * L
    workdir = os.environ.get("RELATE_BENCH_TMP")
    if not workdir and os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * need_est + (16 << 30):
        workdir = "/dev/shm"
    tmp = tempfile.mkdtemp(prefix=f"relate_{name}_", dir=workdir)
    os.environ["RP_IO_THREADS"] = str(os.cpu_count() or 8)   # this leg owns the host: the other ranks are parked
    try:
        out_dir = os.path.join(tmp, "o")
        t0 = time.perf_counter()
        hap, bp, rpos, wb = synth.make_chunk_dir(out_dir, N, L, cfg["seed"], memory_gb=cfg["memory"])
        W = len(wb) - 1
        t_gen = time.perf_counter() - t0
        need = 2.2 * W * N * N * 4
        free_disk = shutil.disk_usage(tmp).free
        if free_disk < need + (8 << 30):
            return {"config": name, "skipped": f"needs {need/1e9:.0f} GB of paint files, {free_disk/1e9:.0f} GB free in {tmp}"}
        files = [os.path.join(out_dir, "chunk_0", "paint", f"relate_{w}.bin") for w in range(W)]

        def run(devs):
            shutil.rmtree(os.path.join(out_dir, "chunk_0"), ignore_errors=True)
            t0 = time.perf_counter()
            st = capi.paint_chunk(out_dir, 0, PAINTING, devices=devs)
            return time.perf_counter() - t0, st
        runs = [run(devices) for _ in range(1 + reps)]            # first call: allocations, pinning (reported apart)
        warm = sorted(t for t, _ in runs[1:])
        t_stage = warm[len(warm) // 2]
        st = runs[-1][1]
        per_dev = capi.stage_device_stats(len(devices))
        md5_n, nbytes = _md5_files(files)
        res = {"config": f"{name}: synthetic block-Kingman N={N} x L={L}, one chunk, --memory {cfg['memory']:g} (W={W} windows, "
                         f"U={st['sites']} visited sites), --painting {PAINTING}; targets sharded over {len(devices)} GPU(s) by "
                         f"rp_paint_chunk (equal-count batches pulled from one counter), chunk files -> paint files",
               "n_devices": len(devices), "scaling": "strong", "workdir": os.path.dirname(tmp) or tmp,
               "ms_stage": 1e3 * t_stage, "ms_stage_runs": [round(1e3 * t, 1) for t, _ in runs], "ms_stage_first_call": 1e3 * runs[0][0],
               "cells_per_s": cells / t_stage, "ms_paint_max": st["ms_paint"], "kernel_cells_per_s": cells / (st["ms_paint"] * 1e-3),
               "kernel": f"paint_kernel<float,{st['words_per_thread']},multi> teams of {st['team_threads']} threads",
               "kernel_frac_nominal": [7.0 * N * d["sites"] / (d["ms_paint"] * 1e-3) / nominal if d["ms_paint"] > 0 else None for d in per_dev],
               "per_device": [{k: d[k] for k in ("n_targets", "sites", "ms_prep", "ms_paint", "ms_rle", "ms_d2h", "launches")} for d in per_dev],
               "breakdown_ms": {k: st[k] for k in ("ms_load", "ms_h2d", "ms_prep", "ms_paint", "ms_rle", "ms_d2h", "ms_write", "ms_total")},
               "paint_file_bytes": nbytes, "files_md5": md5_n, "ms_generate_input": 1e3 * t_gen}
        if len(devices) > 1:                                       # the N-device files must be the 1-device files
            t1, st1 = run(devices[:1])
            md5_1, _ = _md5_files(files)
            res["files_md5_1gpu"] = md5_1
            res["files_identical_to_1gpu"] = (md5_1 == md5_n)
            res["ms_stage_1gpu_same_box"] = 1e3 * t1
        # d_ij through the oracle's one-row lens (oracle/lens.py): GPU paint files vs the fp64 oracle's stepping stones
        r = chunkio.r_from_rpos(rpos)
        theta = float(np.float32(PAINTING.split(",")[0]))
        worst, rows_checked, checked, ndiff, nent, worst_vec = 0.0, 0, [], 0, 0, 0.0
        for w in sorted({0, W // 2, W - 1}):
            idx = lens.paint_file_index(files[w], N)
            lo, hi = int(wb[w]), (int(wb[w + 1]) - 1 if w < W - 1 else L - 1)
            snps = sorted({lo, lo + (hi - lo) // 3, lo + 2 * (hi - lo) // 3, hi})
            for n in (3 + 37 * w) % N, (N // 2 + 11 * w) % N:
                a, sa, la, b, sb, lb = lens.read_target_records(files[w], N, n, idx)
                got = lens.dij_rows(hap, r, rpos, wb, theta, w, n, snps, a, b, sa, sb, la, lb)
                o = oracle.paint_targets(hap, r, wb, theta, n, n + 1)
                assert sa == o["site_begin"][0, w] and sb == o["site_end"][0, w], "boundary SNPs differ from the oracle's"
                oa, ob = lens.collapse(o["alpha"][0, w]), lens.collapse(o["beta"][0, w])
                want = lens.dij_rows(hap, r, rpos, wb, theta, w, n, snps, oa, ob, sa, sb, o["ls_alpha"][0, w], o["ls_beta"][0, w])
                worst = max(worst, float(np.abs(got.astype(np.float64) - want).max()))
                ndiff += int((got != want).sum())
                nent += got.size
                for x, y in ((a, oa), (b, ob)):   # the decoded stepping stones themselves (codec tolerance 1e-3)
                    nz = y != 0
                    worst_vec = max(worst_vec, float((np.abs(x[nz].astype(np.float64) - y[nz]) / y[nz]).max()))
                rows_checked += len(snps)
            checked.append(w)
        tol = 1e-4 * abs(np.log(theta / (1 - theta)))
        res["dij"] = {"windows": checked, "rows": rows_checked, "worst_abs_diff": worst, "tolerance": float(tol), "ok": bool(worst <= tol),
                      "entries": nent, "entries_not_bit_identical": ndiff, "worst_rel_diff_decoded_stepping_stones": worst_vec,
                      "lens": "oracle/lens.py: ro_repaint_section + ro_matrix_row (pinned bit-identical to oracle/_ref/dlens) on the "
                              "decoded GPU records vs the fp64 oracle's stepping stones after the codec's collapse"}
        if name == "config4":   # VERDICT r01 task 5: Paint + window-open wall on ONE GPU without paint files
            try:
                shutil.rmtree(os.path.join(out_dir, "chunk_0"), ignore_errors=True)
                res["e2e_resident_1gpu"] = resident_leg(out_dir, devices[0], N, reps=3)
            except Exception as e:
                res["e2e_resident_1gpu"] = {"error": f"{type(e).__name__}: {e}"}
        return res
    finally:
        os.environ.pop("RP_IO_THREADS", None)
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------------------------------
def tree_builder_leg(device, sizes=((1000, 3), (5000, 2), (10000, 2))):
    """Row f4: rp_minmatch_quickbuild (one CTA per tree) on seeded GetMatrix-shaped matrix sequences (tests/mm_cases.py: the
    first tree without a prior, the next with the prior BuildTopology derives from the previous tree) against the reference's
    own MinMatch::QuickBuild timed on one host core (oracle/_ref/qblens; the oracle port if the lens did not travel).  Checks
    that the GPU merge lists are identical to the CPU ones."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mm_cases
    from oracle import oracle
    from relate_b200 import capi
    out = []
    for N, T in sizes:
        res = []
        with capi.MinMatch(N, mm_cases.THETA, device=device) as g:
            def build(d, prior):
                t0 = time.perf_counter()
                m, st = g.quickbuild(d, prior)
                res.append((m, st, time.perf_counter() - t0))
                return m
            trees = mm_cases.tree_sequence(1, N, "tree", T, oracle.prior_from_merges, build)
            # warm second pass over the same inputs on a fresh handle (the first launch of a process loads the module)
        warm = []
        with capi.MinMatch(N, mm_cases.THETA, device=device) as g:
            for d, prior in trees:
                t0 = time.perf_counter()
                m, st = g.quickbuild(d, prior)
                warm.append((m, st, time.perf_counter() - t0))
        tmpd = tempfile.mkdtemp(prefix="relate_qb_")
        try:
            if os.access(oracle.REF_QBLENS, os.X_OK):
                ref, secs = oracle.reference_quickbuild(N, mm_cases.THETA, trees, tmpd)
                kind = "reference"
            else:
                o = oracle.MinMatchOracle(N, mm_cases.THETA)
                t0 = time.perf_counter()
                ref = [o.quickbuild(d, prior)[0] for d, prior in trees]
                secs, kind = time.perf_counter() - t0, "port"
        finally:
            shutil.rmtree(tmpd, ignore_errors=True)
        same = all(np.array_equal(ref[t], warm[t][0]) and np.array_equal(ref[t], res[t][0]) for t in range(T))
        # a tree occupies one SM: K handles (K windows in BuildTopology's terms) driven from K host threads side by side, the
        # matrices already on the device (rp_minmatch_quickbuild_device: what a consumer that keeps GetMatrix's output in HBM calls)
        conc = None
        if N <= 1000:
            import torch
            dev_trees = [(torch.from_numpy(d).cuda(device), None if prior is None else torch.from_numpy(prior).cuda(device)) for d, prior in trees]
            torch.cuda.synchronize(device)
            K = 64
            handles = [capi.MinMatch(N, mm_cases.THETA, device=device) for _ in range(K)]
            outs = [None] * K

            def work(i):
                outs[i] = [handles[i].quickbuild_device(d.data_ptr(), None if prior is None else prior.data_ptr())[0] for d, prior in dev_trees]
            for rep in range(2):   # (the second round is the warm one)
                ts = [threading.Thread(target=work, args=(i,)) for i in range(K)]
                t0 = time.perf_counter()
                [x.start() for x in ts]
                [x.join() for x in ts]
                wall = time.perf_counter() - t0
            for h in handles:
                h.close()
            conc = {"handles": K, "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "trees_per_s": K * T / wall, "one_handle_trees_per_s": T / sum(w[2] for w in warm),
                    "identical": bool(all(np.array_equal(outs[i][t], ref[t]) for i in range(K) for t in range(T)))}
        out.append({"N": N, "trees": T, "gpu_ms_per_tree_kernel": sum(w[1]["ms_kernel"] for w in warm) / T,
                    "gpu_ms_per_tree_call": 1e3 * sum(w[2] for w in warm) / T, "cpu_ms_per_tree": 1e3 * secs / T, "cpu_kind": kind,
                    "cpu_cores": 1, "draws_per_tree": sum(w[1]["draws"] for w in warm) / T,
                    "general_steps": sum(w[1]["general_steps"] for w in warm), "medium_steps": sum(w[1]["medium_steps"] for w in warm),
                    "merge_lists_identical": bool(same), "concurrent": conc})
    return {"what": "MinMatch::QuickBuild (src/tree_builder.cpp:1060-1303, 2357-2646) per tree, host matrices in (H2D inside the call), "
                    "merge list out; one CTA per tree", "sizes": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling leg (one config-3 chunk over all GPUs)")
    ap.add_argument("--words-per-thread", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from relate_b200 import capi, chunkio, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the painting path has no CPU fallback")
    # stdout carries exactly one JSON line: native libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = sharding.cpu_barrier_group()     # ranks park here (no spinning GPU kernel) during rank 0's sharded leg

    tmp = tempfile.mkdtemp(prefix=f"relate_bench_r{rank}_")
    try:
        out_dir = os.path.join(tmp, "o")
        hap, bp, rpos, wb = make_chunk(out_dir, seed=1 + rank)
        r = chunkio.r_from_rpos(rpos)
        theta = float(np.float32(PAINTING.split(",")[0]))
        W = len(wb) - 1
        cells_per_step = N_HAP * N_HAP * N_SNP

        stream = torch.cuda.Stream(device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
        chunk = capi.DeviceChunk.from_arrays(hap, r, wb, theta, device=local_rank)
        chunk._wb = wb
        chunk.set_tune(words_per_thread=args.words_per_thread, ctas_per_sm=args.ctas_per_sm)
        chunk.set_stream(stream.cuda_stream)

        def step():
            return chunk.paint_targets_device(0, N_HAP)

        sampler = ClockSampler(local_rank)
        sampler.start()
        with torch.cuda.stream(stream):
            for _ in range(args.warmup):
                flush.fill_(1)
                st = step()
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            t_wall0 = time.perf_counter()
            paint_ms, prep_ms, launches = [], [], 0
            for i in range(args.steps):
                flush.fill_(i & 0xFF)          # L2 flush between timed iterations (outside the event pair)
                ev[i][0].record(stream)
                st = step()                   # prep kernels + paint kernel on `stream`
                ev[i][1].record(stream)
                paint_ms.append(st["ms_paint"]); prep_ms.append(st["ms_prep"]); launches += st["launches"]
            torch.cuda.synchronize(dev)
            t_wall = time.perf_counter() - t_wall0
            if world > 1:
                dist.barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        my_ms = sum(step_ms) / len(step_ms)
        U = st["sites"]
        kernel_ms = sum(paint_ms) / len(paint_ms)

        # ---- end to end through the reference-facing stage (files in -> files out) ------------------------
        def e2e_step():
            shutil.rmtree(os.path.join(out_dir, "chunk_0"), ignore_errors=True)
            t0 = time.perf_counter()
            s = capi.paint_chunk(out_dir, 0, PAINTING, devices=[local_rank])
            return time.perf_counter() - t0, s

        for _ in range(max(3, args.warmup)):  # warm-up: pinned rings, device workspaces, page cache of the outputs' directory
            e2e_step()
        if world > 1:
            dist.barrier()
        e2e_runs = [e2e_step() for _ in range(max(3, min(args.steps, 5)))]
        # the timed region is tens of milliseconds; keep the same kernels running (untimed) until nvidia-smi has
        # delivered enough samples for a meaningful "under load" clock reading
        t_extra = time.perf_counter()
        while len(sampler.rows) < 12 and time.perf_counter() - t_extra < 4.0:
            step()
        clocks = sampler.stop()
        my_e2e = sum(t for t, _ in e2e_runs) / len(e2e_runs)
        es = e2e_runs[-1][1]

        (t_max, e2e_max, kern_max) = sharding.allreduce_scalars([my_ms, my_e2e, kernel_ms], "max", device=dev)
        (u_sum,) = sharding.allreduce_scalars([float(U)], "sum", device=dev)

        # ---- strong scaling of one chunk over all GPUs of the job (rank 0 drives every device) ----
        sharded = {}
        if not args.no_sharded:
            chunk.set_stream(None)
            torch.cuda.synchronize(dev)
            if cpu_group is not None:
                dist.barrier(group=cpu_group)            # every rank's own legs are done: the GPUs are idle
            if rank == 0:
                peaks0, _ = measured_peaks()
                legs = ["config3"]
                if world >= 8 or os.environ.get("RELATE_BENCH_CONFIG4") == "1":
                    legs.append("config4")
                for name in legs:
                    try:
                        sharded[name] = sharded_leg(name, list(range(world)), peaks0)
                    except Exception as e:  # the headline line must survive a failure here
                        sharded[name] = {"config": name, "error": f"{type(e).__name__}: {e}"}
            if cpu_group is not None:
                dist.barrier(group=cpu_group)

        if rank == 0:
            peaks, peak_src = measured_peaks()
            mix, scalar = capi.peak_fp32(local_rank)
            nominal = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6
            alg_ops = 7.0 * N_HAP * U                      # 7 FP32 ops per (target, site, reference) pair
            achieved = alg_ops / (kernel_ms * 1e-3)
            hbm_alg = (N_SNP * chunk.words_per_snp * 4) + 8.0 * U + 2.0 * W * N_HAP * N_HAP * 4   # bit matrix + tables + stones
            line = {
                "metric": METRIC, "value": world * cells_per_step / (t_max * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_max, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": CONFIG,
                "run": {"windows": W, "visited_sites": U, "chunks": "one such chunk per GPU (weak scaling)",
                        "team_threads": st["team_threads"], "words_per_thread": st["words_per_thread"], "ctas": st["ctas"]},
                "roofline": {"bound": "fp32", "achieved": achieved / 1e12, "peak": mix / 1e12, "unit": "TFLOP/s",
                             "frac": achieved / mix, "traffic": (ncu_traffic() or (None, None))[0],
                             "traffic_source": (ncu_traffic() or (None, None))[1],
                             "kernel": "paint_kernel", "kernel_ms": kernel_ms, "prep_ms": sum(prep_ms) / len(prep_ms),
                             "algorithmic_ops": alg_ops, "executed_fp32_lane_ops": 6.0 * N_HAP * U,
                             "peak_source": "measured on this box: rp_peak_fp32 (packed add + predicated-mul mix, no memory)",
                             "peak_scalar_mix": scalar / 1e12, "peak_nominal_issue": nominal / 1e12,
                             "frac_of_nominal_issue": achieved / nominal,
                             "hbm": {"algorithmic_bytes": hbm_alg, "achieved_gbs": hbm_alg / (kernel_ms * 1e-3) / 1e9,
                                     "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src}},
                "e2e": {"value": world * cells_per_step / e2e_max, "unit": UNIT,
                        "h2d_bytes_per_step": es["h2d_bytes"], "d2h_bytes_per_step": es["d2h_bytes"],
                        "seconds_per_step": e2e_max, "runs_ms": [round(1e3 * t, 3) for t, _ in e2e_runs],
                        "call": "rp_paint_chunk (== relate --mode Paint): chunk files -> paint files",
                        "breakdown_ms": {k: es[k] for k in ("ms_load", "ms_h2d", "ms_prep", "ms_paint", "ms_rle", "ms_d2h", "ms_write", "ms_total")}},
                "gpu_launches": launches,
                "clocks": clocks,
                "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            }
            if "config3" in sharded:
                line["sharded"] = sharded["config3"]
            if "config4" in sharded:
                line["sharded_config4"] = sharded["config4"]
            if world == 1:
                line["window_repaint"] = window_repaint_measure(chunk, rpos, W, peaks)
                try:
                    line["e2e_resident"] = dict(resident_leg(out_dir, local_rank, N_HAP), workload=WORKLOAD)
                except Exception as e:
                    line["e2e_resident"] = {"error": f"{type(e).__name__}: {e}"}
            if world == 1:
                try:
                    line["tree_builder"] = tree_builder_leg(local_rank)
                except Exception as e:
                    line["tree_builder"] = {"error": f"{type(e).__name__}: {e}"}
            if world == 1 and not args.no_cpu_baseline:
                cb = cpu_baseline_sample()
                line["cpu_baseline"] = cb
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)
        chunk.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
