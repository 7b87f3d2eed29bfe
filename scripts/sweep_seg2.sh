python scripts/prof_case.py 1000 50000 5 0 0 1000 0 0 2>&1 | tail -1 | cut -c1-150
python scripts/prof_case.py 2000 20000 3 0 0 2000 0 0 2>&1 | tail -1| cut -c1-150
python scripts/prof_case.py 5000 20000 2 0 0 5000 0 0 2>&1 | tail -1|  cut -c1-150
python scripts/prof_case.py 10000 20000 2 0 0 2236 0 0 2>&1 | tail -1| cut -c1-150
