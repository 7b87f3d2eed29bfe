for rep in 1 2 3; do python scripts/prof_case.py 1000 50000 8 0 0 1000 2>&1 | tail -1 | cut -c1-140; done
python scripts/prof_case.py 2000 20000 5 0 0 2000 2>&1 | tail -1 | cut -c1-140
python scripts/prof_case.py 5000 20000 2 0 0 5000 2>&1 | tail -1 | cut -c1-140
