"""Paint one synthetic chunk a few times (for ncu / quick timing).  usage: prof_case.py N L [reps] [wpt] [ctas_per_sm] [nk] [nodense] [segments] [cluster]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import synth, chunkio, capi
N, L = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wpt = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
nk = int(sys.argv[6]) if len(sys.argv) > 6 else N
nodense = int(sys.argv[7]) if len(sys.argv) > 7 else 0
segments = int(sys.argv[8]) if len(sys.argv) > 8 else 0
cluster = int(sys.argv[9]) if len(sys.argv) > 9 else 0
hap, bp = synth.block_kingman(N, L, 1)
r = chunkio.r_from_rpos(chunkio.uniform_map_rpos(bp))
mem = 5.0 if N <= 1000 else (50.0 if N <= 5000 else 100.0)
wb = chunkio.window_boundaries(hap, mem)
with capi.DeviceChunk.from_arrays(hap, r, wb, 0.001) as c:
    import ctypes as C
    t = capi.RpTune(wpt, cps); t.reserved[0] = nodense; t.reserved[1] = segments; t.reserved[3] = cluster
    capi.check(capi.lib().rp_chunk_set_tune(c._h, C.byref(t)))
    for it in range(reps):
        st = c.paint_targets_device(0, nk)
    U, t = st['sites'], st['ms_paint'] * 1e-3
    print(f"N={N} L={L} W={len(wb)-1} nk={nk} seg={segments} T={st['team_threads']} wpt={st['words_per_thread']} ctas={st['ctas']} paint_ms={st['ms_paint']:.3f} prep_ms={st['ms_prep']:.3f} "
          f"U={U} cells/s={nk*N*L/t:.3e} frac7={7*N*U/t/37.2e12:.3f}")
