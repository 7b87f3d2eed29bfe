nproc
for w in 8 16 24 48; do echo "RP_WRITERS=$w"; RP_WRITERS=$w RP_IO_THREADS=64 python scripts/prof_stage.py 5000 100000 50 4 | tail -2; done
