for r in 4 8 16; do for k in 1024 4096 16384; do echo "readers=$r slice_kb=$k"; RP_READERS=$r RP_SLICE_KB=$k python scripts/prof_stage.py 1000 50000 5 6 | tail -2; done; done
