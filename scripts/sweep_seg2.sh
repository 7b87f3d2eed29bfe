for rep in 1 2; do
for v in ns20 ns256 ns1000; do for seg in 8; do RELATE_PAINT_LIB=$PWD/variants/lib_$v.so python scripts/prof_case.py 1000 50000 5 0 0 1000 0 $seg 2>&1 | tail -1 | sed "s/^/$v /" | cut -c1-120; done; done
python scripts/prof_case.py 1000 50000 5 0 0 1000 0 0 2>&1 | tail -1 | sed "s/^/ns64auto /" | cut -c1-120
python scripts/prof_case.py 1000 50000 5 0 0 1000 0 1 2>&1 | tail -1 | sed "s/^/seg1 /" | cut -c1-120
done
python scripts/prof_case.py 5000 8000 2 0 0 5000 0 0 2>&1 | tail -1| cut -c1-150
python scripts/prof_case.py 5000 20000 2 0 0 5000 0 0 2>&1 | tail -1| cut -c1-150
python scripts/prof_case.py 5000 20000 2 0 0 5000 0 1 2>&1 | tail -1| cut -c1-150
python scripts/prof_case.py 10000 20000 2 0 0 2236 0 0 2>&1 | tail -1| cut -c1-150
python scripts/prof_case.py 10000 20000 2 0 0 2236 0 1 2>&1 | tail -1| cut -c1-150
