for ck in 2 3 4; do echo "CK=$ck"; RP_REPAINT_CK=$ck python scripts/prof_window.py 1000 50000 | tail -1; RP_REPAINT_CK=$ck python scripts/prof_window.py 600 30000 | tail -1; done
