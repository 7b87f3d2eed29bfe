nproc; df -h /dev/shm | tail -1
for w in 8 16 32; do echo "shm RP_WRITERS=$w"; RELATE_TMP=/dev/shm RP_WRITERS=$w RP_IO_THREADS=64 python scripts/prof_stage.py 5000 100000 50 4 | tail -2; done
for w in 8 16; do echo "tmp RP_WRITERS=$w"; RP_WRITERS=$w RP_IO_THREADS=64 python scripts/prof_stage.py 5000 100000 50 3 | tail -1; done
