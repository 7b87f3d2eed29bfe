"""Stage timing of rp_paint_chunk on a synthetic chunk, first (cold: allocations, pinning) and later (warm) calls.
usage: prof_stage.py N L memory_gb [calls]"""
import os, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relate_b200 import synth, capi
N, L, mem = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 3
tmp = tempfile.mkdtemp(prefix="relate_stage_", dir=os.environ.get("RELATE_TMP"))
try:
    synth.make_chunk_dir(os.path.join(tmp, "o"), N, L, seed=2, memory_gb=mem)
    ndev = capi.lib().rp_device_count()
    for i in range(calls):
        shutil.rmtree(os.path.join(tmp, "o", "chunk_0"), ignore_errors=True)
        t0 = time.perf_counter()
        st = capi.paint_chunk(os.path.join(tmp, "o"), 0, "0.001,1", devices=list(range(ndev)))
        dt = time.perf_counter() - t0
        keys = ("ms_load", "ms_h2d", "ms_prep", "ms_paint", "ms_rle", "ms_d2h", "ms_write", "ms_total")
        print(f"call {i}: wall {dt*1e3:.1f} ms  " + " ".join(f"{k[3:]}={st[k]:.1f}" for k in keys) + f" launches={st['launches']}", flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
