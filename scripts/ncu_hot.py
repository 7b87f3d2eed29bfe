"""Per-instruction stall samples of the hottest region of an exported ncu source page:
   ncu -i rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    n = int(r[col["# Samples"]]); tot += n
print("total samples", tot)
for i, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    n = int(r[col["# Samples"]])
    ex = int(r[col["Instructions Executed"]])
    if n < mins: continue
    top = sorted(((int(r[col[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {ex:9d} {n:6d}  {r[col['Source']].strip():60s} " + " ".join(f"{s}:{v}" for v, s in top if v))
