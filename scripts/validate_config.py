"""Whole-stage run at a BASELINE.json config shape (run under gpurun, any number of GPUs):
`relate --mode Paint` (CLI, all visible GPUs, targets sharded) on N x L synthetic haplotypes, stage breakdown from the
CLI banner, then a few records decoded and compared with the oracle.  usage: validate_config.py N L memory_gb [ntargets_checked]"""
import os, subprocess, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import synth, chunkio, capi
from oracle import oracle

N, L, mem = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
ncheck = int(sys.argv[4]) if len(sys.argv) > 4 else 3
tmp = tempfile.mkdtemp(prefix="relate_cfg_", dir=os.environ.get("RELATE_TMP"))
try:
    t0 = time.time()
    hap, bp, rpos, wb = synth.make_chunk_dir(os.path.join(tmp, "o"), N, L, seed=2, memory_gb=mem)
    W = len(wb) - 1
    print(f"generated N={N} L={L} W={W} in {time.time()-t0:.1f}s; devices: {capi.lib().rp_device_count()}", flush=True)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "relate_b200", "bin", "relate")
    for rep in range(2):
        shutil.rmtree(os.path.join(tmp, "o", "chunk_0"), ignore_errors=True)
        t0 = time.time()
        p = subprocess.run([exe, "--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", "0.001,1"], cwd=tmp,
                           capture_output=True, text=True)
        dt = time.time() - t0
        banner = [l for l in p.stderr.splitlines() if l.startswith("GPU Paint")]
        print(f"run {rep}: rc {p.returncode} wall {dt:.2f}s  cells/s {N*N*L/dt:.3e} | {banner}", flush=True)
        assert p.returncode == 0, p.stderr
    sizes = [os.path.getsize(os.path.join(tmp, "o", "chunk_0", "paint", f"relate_{w}.bin")) for w in range(W)]
    print("paint file bytes", sum(sizes), "vs raw stepping stones", 2 * W * N * N * 4)
    r = chunkio.r_from_rpos(rpos)
    theta = float(np.float32(0.001))
    worst = 0.0
    ks = sorted(set(int(x) for x in np.linspace(0, N - 1, ncheck)))
    ora = {k: oracle.paint_targets(hap, r, wb, theta, k, k + 1) for k in ks}
    for w in range(0, W, max(1, W // 4)):
        recs = chunkio.read_paint_file(os.path.join(tmp, "o", "chunk_0", "paint", f"relate_{w}.bin"), N)
        assert len(recs) == N
        for k in ks:
            o = ora[k]
            a0, b0, ra, rb = recs[k]
            assert (a0, b0) == (wb[w], wb[w + 1] - 1)
            assert ra.site == o["site_begin"][0, w] and rb.site == o["site_end"][0, w]
            for rec, pre in ((ra, o["alpha"][0, w]), (rb, o["beta"][0, w])):
                dec = rec.expand().astype(np.float64); pre = pre.astype(np.float64)
                msk = pre != 0
                worst = max(worst, float((np.abs(dec[msk] - pre[msk]) / pre[msk]).max()))
    print("worst decoded-vs-oracle relative difference (codec tolerance 1e-3):", worst)
    assert worst < 1.2e-3
    print("OK")
finally:
    shutil.rmtree(tmp, ignore_errors=True)
