python scripts/prof_case.py 5000 20000 2 1 0 5000 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 5000 20000 2 2 0 5000 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 10000 20000 2 1 0 2236 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 10000 20000 2 2 0 2236 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 3000 20000 2 1 0 3000 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 3000 20000 2 2 0 3000 2>&1 | tail -1 | cut -c1-160
