"""Samples and executed warp instructions per CUDA source line: joins an ncu SASS source page with nvdisasm line info.
   nvdisasm --print-line-info -c x.cubin > x.dis ; ncu -i rep --page source --csv > src.csv
   python scripts/ncu_by_line.py src.csv x.dis <kernel-substring> [top]"""
import csv, re, sys, collections
src, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
addr2line, cur, on = {}, None, False
for ln in open(dis, errors="replace"):
    if ln.startswith("//---") and ".text." in ln:
        on = kern in ln
    if not on: continue
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m: cur = int(m.group(1)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None: base = a
    line = addr2line.get(a - base)
    e = agg[line]
    e[0] += int(r[col["# Samples"]]); e[1] += int(r[col["Instructions Executed"]])
    for s in stalls: e[2][s[6:]] += int(r[col[s]])
tot_s = sum(e[0] for e in agg.values()); tot_i = sum(e[1] for e in agg.values())
print("samples", tot_s, "warp instructions", tot_i)
lines = open(re.search(r'File "(.*?)"', open(dis, errors="replace").read()).group(1)).read().split("\n")
for line, e in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    txt = lines[line - 1].strip()[:70] if line else "?"
    print(f"{line!s:>5} {100*e[0]/tot_s:5.1f}% smp {100*e[1]/tot_i:5.1f}% ins  {txt:70s} " + " ".join(f"{k}:{v}" for k, v in e[2].most_common(3)))
