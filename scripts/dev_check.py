"""Developer scratch check: CUDA path vs oracle on small synthetic chunks (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from relate_b200 import synth, chunkio, capi
from oracle import oracle

def case(N, L, W, seed, fp64, theta=0.001, wpt=0, nk=None):
    hap, bp = synth.block_kingman(N, L, seed)
    rpos = chunkio.uniform_map_rpos(bp)
    r = chunkio.r_from_rpos(rpos)
    wb = np.linspace(0, L, W + 1).astype(np.int32); wb[0] = 0; wb[-1] = L
    nk = N if nk is None else nk
    with capi.DeviceChunk.from_arrays(hap, r, wb, theta, fp64=fp64) as c:
        if wpt: c.set_tune(words_per_thread=wpt)
        g = c.paint_targets(0, nk)
    o = oracle.paint_targets(hap, r, wb, theta, 0, nk)
    ok_sites = np.array_equal(g.site_begin, o["site_begin"]) and np.array_equal(g.site_end, o["site_end"])
    def rel(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        den = np.maximum(np.abs(b.astype(np.float64)), 1e-300)
        m = (b != 0) | (a != 0)
        return float((d[m] / den[m]).max()) if m.any() else 0.0
    ra, rb = rel(g.alpha, o["alpha"]), rel(g.beta, o["beta"])
    la = float(np.abs(g.ls_alpha - o["ls_alpha"]).max()); lb = float(np.abs(g.ls_beta - o["ls_beta"]).max())
    exact = np.array_equal(g.alpha, o["alpha"]) and np.array_equal(g.beta, o["beta"]) and np.array_equal(g.ls_alpha, o["ls_alpha"]) and np.array_equal(g.ls_beta, o["ls_beta"])
    print(f"N={N} L={L} W={W} fp64={fp64} wpt={g.stats['words_per_thread']} T={g.stats['team_threads']} ctas={g.stats['ctas']} sites_ok={ok_sites} "
          f"rel_alpha={ra:.3g} rel_beta={rb:.3g} dls_a={la:.3g} dls_b={lb:.3g} exact={exact} paint_ms={g.stats['ms_paint']:.3f} prep_ms={g.stats['ms_prep']:.3f}", flush=True)
    return g, o

if __name__ == "__main__":
    print(capi.lib().rp_version().decode(), "devices", capi.lib().rp_device_count())
    x = np.concatenate([np.float32(10.0) ** np.arange(-30, 30, dtype=np.float32), np.random.default_rng(0).random(1000, dtype=np.float32) * 1e-9])
    print("fast_log exact:", np.array_equal(capi.fast_log_device(x), oracle.fast_log(x)))
    case(64, 500, 3, 1, True)
    case(64, 500, 3, 1, False)
    case(200, 2000, 5, 2, True)
    case(200, 2000, 5, 2, False)
    case(8, 3000, 4, 3, True)
    case(8, 3000, 4, 3, False)
    case(1000, 2000, 4, 4, True, nk=64)
    case(1000, 2000, 4, 4, False, nk=64)
    case(1500, 1500, 3, 5, False, nk=48)             # WPT=2 single warp
    case(1500, 1500, 3, 5, False, nk=48, wpt=1)      # multi warp WPT=1
    case(1500, 1500, 3, 5, True, nk=48)              # fp64 multi
    case(2500, 1200, 3, 6, False, nk=32)             # multi WPT=2
    case(5000, 600, 2, 7, False, nk=16)
    case(5000, 600, 2, 7, False, nk=16, wpt=1)
    # timing at config-2 shape
    N, L = 1000, 50000
    hap, bp = synth.block_kingman(N, L, 1)
    rpos = chunkio.uniform_map_rpos(bp); r = chunkio.r_from_rpos(rpos)
    wb = chunkio.window_boundaries(hap, 5.0)
    print("config2 W =", len(wb) - 1)
    with capi.DeviceChunk.from_arrays(hap, r, wb, 0.001) as c:
        for it in range(4):
            st = c.paint_targets_device(0, N)
            print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k != 'dev_ptrs'}, flush=True)
        U = st['sites']
        t = st['ms_paint'] * 1e-3
        print(f"U={U} cells/s={N*N*L/t:.3e} fp32op/s(7NU)={7*N*U/t:.3e} frac_of_37.2T={7*N*U/t/37.2e12:.3f}")
