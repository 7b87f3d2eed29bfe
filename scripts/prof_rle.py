"""Times the device record encoder (rle_kernel count + emit passes) on a synthetic chunk.  usage: prof_rle.py N L nk [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from relate_b200 import synth, chunkio, capi
N, L, nk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
hap, bp = synth.block_kingman(N, L, 1)
r = chunkio.r_from_rpos(chunkio.uniform_map_rpos(bp))
mem = 5.0 if N <= 1000 else (50.0 if N <= 5000 else 100.0)
wb = chunkio.window_boundaries(hap, mem)
with capi.DeviceChunk.from_arrays(hap, r, wb, 0.001) as c:
    for it in range(reps):
        off = np.zeros(c.W + 1, np.int64)
        st = capi.RpStats()
        capi.check(capi.lib().rp_paint_records(c._h, 0, nk, capi._ptr(off), C.byref(st)))
        print(f"N={N} L={L} W={c.W} nk={nk} paint_ms={st.ms_paint:.3f} rle_ms={st.ms_rle:.3f} image_bytes={int(off[-1])} raw={2*nk*c.W*N*4} "
              f"runs/elem={(int(off[-1]) - 64*nk*c.W)/8/(2*nk*c.W*N):.3f}")
