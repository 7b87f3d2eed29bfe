for cps in 0 14 15; do for seg in 1 4 8 16 32; do python scripts/prof_case.py 1000 50000 3 0 $cps 1000 0 $seg 2>&1 | tail -1 | sed "s/^/cps=$cps /" | cut -c1-150; done; done
for seg in 1 8 16; do python scripts/prof_case.py 2000 20000 3 0 0 2000 0 $seg 2>&1 | tail -1| cut -c1-150; done
for seg in 1 8; do python scripts/prof_case.py 5000 8000 2 0 0 5000 0 $seg 2>&1 | tail -1| cut -c1-150; done
for seg in 1 8; do python scripts/prof_case.py 10000 4000 2 0 0 2236 0 $seg 2>&1 | tail -1| cut -c1-150; done
