"""A/B of the look-ahead painter against the plain one (single-warp teams): kernel time and agreement.
usage: ab_look.py [N L]..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import capi, chunkio, synth
from oracle import oracle

shapes = [(1000, 50000), (2000, 20000), (600, 20000), (200, 20000)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
theta = float(np.float32(0.001))
for N, L in shapes:
    hap, bp = synth.block_kingman(N, L, 1)
    r = chunkio.r_from_rpos(chunkio.uniform_map_rpos(bp))
    wb = chunkio.window_boundaries(hap, 5.0)
    with capi.DeviceChunk.from_arrays(hap, r, wb, theta) as c:
        res = {}
        for name, plain in (("look", False), ("plain", True)):
            c.set_tune(plain_kernel=plain)
            ts = []
            for i in range(6):
                st = c.paint_targets_device(0, N)
                ts.append(st["ms_paint"])
            g = c.paint_targets(0, min(N, 64))
            res[name] = (sorted(ts[1:])[len(ts[1:]) // 2], g, st)
        tl, gl, stl = res["look"]
        tp, gp, stp = res["plain"]
        U = stl["sites"]
        nominal = 148 * 128 * 1.965e9
        def rel(a, b):
            a = a.astype(np.float64); b = b.astype(np.float64); m = (a != 0) | (b != 0)
            return float((np.abs(a - b)[m] / np.maximum(np.abs(b[m]), 1e-300)).max())
        o = oracle.paint_targets(hap, r, wb, theta, 0, 8)
        print(f"N={N} L={L} W={len(wb)-1} U={U}: look {tl:.3f} ms ({7*N*U/(tl*1e-3)/nominal:.3f} of nominal), plain {tp:.3f} ms ({7*N*U/(tp*1e-3)/nominal:.3f}); "
              f"speed-up {tp/tl:.3f}; look vs plain rel {rel(gl.alpha, gp.alpha):.2e}/{rel(gl.beta, gp.beta):.2e}; "
              f"look vs oracle {rel(gl.alpha[:8], o['alpha']):.2e}/{rel(gl.beta[:8], o['beta']):.2e}; plain vs oracle {rel(gp.alpha[:8], o['alpha']):.2e}/{rel(gp.beta[:8], o['beta']):.2e}; "
              f"ls diff {np.abs(gl.ls_alpha[:8]-o['ls_alpha']).max():.2e}/{np.abs(gl.ls_beta[:8]-o['ls_beta']).max():.2e}", flush=True)
