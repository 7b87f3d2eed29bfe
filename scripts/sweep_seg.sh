# chain-segment sweep of the paint kernel: usage under gpurun: bash scripts/sweep_seg.sh
# prof_case.py args: N L reps wpt ctas_per_sm nk nodense segments
for seg in 1 2 4 8 16; do python scripts/prof_case.py 1000 50000 5 0 0 1000 0 $seg 2>&1 | tail -1 | cut -c1-150; done
for seg in 1 0; do python scripts/prof_case.py 2000 20000 3 0 0 2000 0 $seg 2>&1 | tail -1 | cut -c1-150; done
for seg in 1 0; do python scripts/prof_case.py 5000 20000 2 0 0 5000 0 $seg 2>&1 | tail -1 | cut -c1-150; done
for seg in 1 0; do python scripts/prof_case.py 10000 20000 2 0 0 2236 0 $seg 2>&1 | tail -1 | cut -c1-150; done
