for nk in 74 148 296 444 592 740 888 1000; do python scripts/prof_case.py 1000 50000 3 0 0 $nk 2>&1 | tail -1; done
