import sys, os, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mm_cases, torch
from relate_b200 import capi
from oracle import oracle
N = 1000
res = []
o = oracle.MinMatchOracle(N, mm_cases.THETA)
trees = mm_cases.tree_sequence(1, N, "tree", 3, oracle.prior_from_merges, lambda d, p: o.quickbuild(d, p)[0])
dev = [(torch.from_numpy(d).cuda(), None if p is None else torch.from_numpy(p).cuda()) for d, p in trees]
torch.cuda.synchronize()
for K in (1, 4, 16, 64):
    hs = [capi.MinMatch(N, mm_cases.THETA) for _ in range(K)]
    ms = [[] for _ in range(K)]
    def work(i):
        for rep in range(3):
            for d, p in dev:
                m, st = hs[i].quickbuild_device(d.data_ptr(), None if p is None else p.data_ptr()); ms[i].append(st["ms_kernel"])
    for rnd in range(2):
        for m_ in ms: m_.clear()
        ts = [threading.Thread(target=work, args=(i,)) for i in range(K)]
        t0 = time.perf_counter(); [x.start() for x in ts]; [x.join() for x in ts]; wall = time.perf_counter() - t0
    print(f"K={K}: {K*9/wall:.0f} trees/s, mean ms_kernel {np.mean([x for m_ in ms for x in m_]):.2f}")
    for h in hs: h.close()
