"""Accuracy of the fp32 painter against the fp64 oracle on one synthetic chunk: max and 99.99th-percentile relative
error of the stepping stones.  usage: acc_check.py [N] [L] [W]   (RELATE_PAINT_LIB selects a variant build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import synth, chunkio, capi
from oracle import oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 6
hap, bp = synth.block_kingman(N, L, 5)
r = chunkio.r_from_rpos(chunkio.uniform_map_rpos(bp))
wb = np.linspace(0, L, W + 1).astype(np.int32)
theta = float(np.float32(0.001))
with capi.DeviceChunk.from_arrays(hap, r, wb, theta) as c:
    g = c.paint_targets(0, N)
o = oracle.paint_targets(hap, r, wb, theta, 0, N)
for name in ("alpha", "beta"):
    a, b = getattr(g, name).astype(np.float64), o[name].astype(np.float64)
    m = b != 0
    rel = np.abs(a[m] - b[m]) / np.abs(b[m])
    print(f"{name}: max rel {rel.max():.3e}  p99.99 {np.quantile(rel, 0.9999):.3e}  median {np.median(rel):.3e}")
print("ls:", float(np.abs(g.ls_alpha - o["ls_alpha"]).max()), float(np.abs(g.ls_beta - o["ls_beta"]).max()))
