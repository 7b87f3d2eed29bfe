"""One-off validation of the whole stage at a 1000G-scale sample count (run under gpurun):
`relate --mode Paint` (CLI, all visible GPUs) on N=10,000 haplotypes, then decode a few records and compare with
the oracle's pre-RLE vectors.  usage: validate_large.py [L] [ngpu]"""
import os, subprocess, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import synth, chunkio, capi
from oracle import oracle

N = 10000
L = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
tmp = tempfile.mkdtemp(prefix="relate_large_")
try:
    t0 = time.time()
    hap, bp, rpos, wb = synth.make_chunk_dir(os.path.join(tmp, "o"), N, L, seed=3, memory_gb=100.0)
    W = len(wb) - 1
    print(f"generated N={N} L={L} W={W} in {time.time()-t0:.1f}s", flush=True)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "relate_b200", "bin", "relate")
    t0 = time.time()
    p = subprocess.run([exe, "--mode", "Paint", "--chunk_index", "0", "-o", "o", "--painting", "0.001,1"], cwd=tmp,
                       capture_output=True, text=True, env=dict(os.environ, RP_TRACE="1"))
    dt = time.time() - t0
    print("\n".join(p.stderr.strip().splitlines()[-25:]))
    print("rc", p.returncode, f"wall {dt:.2f}s  cells/s {N*N*L/dt:.3e}", flush=True)
    assert p.returncode == 0
    sizes = [os.path.getsize(os.path.join(tmp, "o", "chunk_0", "paint", f"relate_{w}.bin")) for w in range(W)]
    print("file bytes", sum(sizes), "vs raw", 2 * W * N * N * 4)
    r = chunkio.r_from_rpos(rpos)
    theta = float(np.float32(0.001))
    worst = 0.0
    check_w = range(W) if W <= 8 else sorted({0, 1, W // 2, W - 2, W - 1})
    print("checking windows", list(check_w), flush=True)
    for w in check_w:
        recs = chunkio.read_paint_file(os.path.join(tmp, "o", "chunk_0", "paint", f"relate_{w}.bin"), N)
        assert len(recs) == N
        for k in (0, 4999, 9999):
            o = oracle.paint_targets(hap, r, wb, theta, k, k + 1)
            a0, b0, ra, rb = recs[k]
            assert (a0, b0) == (wb[w], wb[w + 1] - 1)
            assert ra.site == o["site_begin"][0, w] and rb.site == o["site_end"][0, w]
            for rec, pre in ((ra, o["alpha"][0, w]), (rb, o["beta"][0, w])):
                dec = rec.expand().astype(np.float64); pre = pre.astype(np.float64)
                m = pre != 0
                worst = max(worst, float((np.abs(dec[m] - pre[m]) / pre[m]).max()))
    print("worst decoded-vs-oracle relative difference (codec tolerance 1e-3):", worst)
    assert worst < 1.2e-3
    print("OK")
finally:
    shutil.rmtree(tmp, ignore_errors=True)
