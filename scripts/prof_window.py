"""Paint all targets of a synthetic chunk, then open one window from the resident stepping stones (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import synth, chunkio, capi
N, L = int(sys.argv[1]), int(sys.argv[2])
hap, bp = synth.block_kingman(N, L, 1)
rpos = chunkio.uniform_map_rpos(bp)
r = chunkio.r_from_rpos(rpos)
wb = chunkio.window_boundaries(hap, 5.0 if N <= 1000 else 50.0)
with capi.DeviceChunk.from_arrays(hap, r, wb, 0.001) as c:
    c.paint_targets_device(0, N)
    for rep in range(2):
        with capi.Window.open_resident(c, (len(wb) - 1) // 2, rpos) as win:
            d = win.distance(int(wb[(len(wb) - 1) // 2]) + 5)
            print("rows", win.rows, "repaint_ms", win.stats["ms_paint"], "d[0,1]", d[0, 1])
