python scripts/prof_case.py 1000 50000 5 0 0 1000 0 0 0 0 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 1000 50000 5 0 0 1000 0 0 0 1 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 1000 50000 5 0 0 500 0 0 0 0 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 1000 50000 5 0 0 296 0 0 0 0 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 2000 20000 5 0 0 1000 0 0 0 0 2>&1 | tail -1 | cut -c1-160
python scripts/prof_case.py 2000 20000 5 0 0 296 0 0 0 0 2>&1 | tail -1 | cut -c1-160
