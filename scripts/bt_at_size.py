"""BuildTopology through Relate_gpu at a given size, trees by the reference's CPU QuickBuild vs by rp_minmatch_quickbuild:
   python scripts/bt_at_size.py N L [stock]
GPU Paint of a synthetic one-window chunk, then `Relate_gpu --mode BuildTopology` twice on the same paint files
(RELATE_GPU_MINMATCH=0 / default) and, with `stock`, the unmodified reference (CPU RePaintSection + GetMatrix + QuickBuild).
Prints wall times, the time inside QuickBuild either way, and whether the .anc/.mut files are byte-identical."""
import filecmp, os, re, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relate_b200 import capi, synth
from oracle import oracle

N, L = int(sys.argv[1]), int(sys.argv[2])
stock = len(sys.argv) > 3 and sys.argv[3] == "stock"
tmp = tempfile.mkdtemp(prefix="relate_bt_", dir=os.environ.get("RELATE_TMP"))
try:
    tags = ["cpu", "gpu"] + (["stock"] if stock else [])
    only_gpu = os.environ.get("RELATE_BT_ONLY_GPU") is not None
    synth.make_chunk_dir(os.path.join(tmp, "cpu", "o"), N, L, seed=31, n_windows=1)
    t0 = time.perf_counter()
    capi.paint_chunk(os.path.join(tmp, "cpu", "o"), 0, "0.001,1")
    print(f"N={N} L={L}: GPU Paint {time.perf_counter() - t0:.2f} s")
    for tag in tags[1:]:
        shutil.copytree(os.path.join(tmp, "cpu"), os.path.join(tmp, tag))
    bt = ["--mode", "BuildTopology", "--chunk_index", "0", "--first_section", "0", "--last_section", "0", "-o", "o",
          "--painting", "0.001,1", "--seed", "1"]
    res = {}
    for tag in (tags if not only_gpu else ["gpu"]):
        exe = oracle.REF_RELATE if tag == "stock" else oracle.REF_RELATE_GPU
        env = dict(os.environ, RELATE_GPU_MINMATCH_STATS="1", RELATE_GPU_MINMATCH_MIN_N="0")
        if tag == "cpu":
            env["RELATE_GPU_MINMATCH"] = "0"
        t0 = time.perf_counter()
        p = subprocess.run([exe] + bt, cwd=os.path.join(tmp, tag), capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr[-2000:]
        if os.environ.get("RELATE_BT_STDERR"):
            print("\n".join(l for l in p.stderr.splitlines() if l.startswith("mm_prof"))[-1500:])
        m = re.search(r"QuickBuild: (\d+) trees on the GPU \((\d+) by the reference's code\), ([0-9.]+) s in the call, ([0-9.]+) s in the kernel", p.stderr)
        res[tag] = (wall, m.groups() if m else None)
        print(f"  {tag:5s}: BuildTopology wall {wall:8.2f} s" + (f"; QuickBuild {m.group(3)} s for {int(m.group(1)) + int(m.group(2))} trees "
              f"({m.group(1)} on the GPU, kernel {m.group(4)} s){p.stderr[m.end():].splitlines()[0] if m else ''}" if m else " (unmodified reference: CPU RePaintSection + GetMatrix + QuickBuild)"))
    if only_gpu:
        sys.exit(0)
    same = all(filecmp.cmp(os.path.join(tmp, "cpu", "o", "chunk_0", f"o_0.{e}"), os.path.join(tmp, "gpu", "o", "chunk_0", f"o_0.{e}"), shallow=False)
               for e in ("anc", "mut"))
    print(f"  .anc/.mut byte-identical between CPU and GPU tree builder: {same}")
    assert same
finally:
    shutil.rmtree(tmp, ignore_errors=True)
