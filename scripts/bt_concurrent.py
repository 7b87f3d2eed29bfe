"""P concurrent `Relate_gpu --mode BuildTopology` processes on one GPU (what RelateParallel.sh does over sections / chunks):
   python scripts/bt_concurrent.py N L P"""
import os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relate_b200 import capi, synth
from oracle import oracle
N, L, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tmp = tempfile.mkdtemp(prefix="relate_btc_")
try:
    synth.make_chunk_dir(os.path.join(tmp, "p0", "o"), N, L, seed=31, n_windows=1)
    capi.paint_chunk(os.path.join(tmp, "p0", "o"), 0, "0.001,1")
    for i in range(1, P):
        shutil.copytree(os.path.join(tmp, "p0"), os.path.join(tmp, f"p{i}"))
    bt = [oracle.REF_RELATE_GPU, "--mode", "BuildTopology", "--chunk_index", "0", "--first_section", "0", "--last_section", "0", "-o", "o",
          "--painting", "0.001,1", "--seed", "1"]
    for procs in ([1, P] if P > 1 else [1]):
        for mode in ("gpu", "cpu"):
            env = dict(os.environ, RELATE_GPU_MINMATCH_MIN_N="0", **({"RELATE_GPU_MINMATCH": "0"} if mode == "cpu" else {}))
            t0 = time.perf_counter()
            ps = [subprocess.Popen(bt, cwd=os.path.join(tmp, f"p{i}"), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(procs)]
            assert all(p.wait() == 0 for p in ps)
            print(f"N={N} L={L}: {procs} process(es), trees on the {mode}: wall {time.perf_counter() - t0:.2f} s", flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
