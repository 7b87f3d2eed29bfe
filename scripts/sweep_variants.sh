#!/bin/bash
# usage: sweep_variants.sh name... ; prints paint_ms for nk in 296 (1 chain/sched), 888 (3), 1000 (3-4) at config 2
for v in "$@"; do
  for nk in 296 888 1000; do
    RELATE_PAINT_LIB=$PWD/variants/lib_$v.so python scripts/prof_case.py 1000 50000 3 0 0 $nk 2>&1 | tail -1 | sed "s/^/$v /" | cut -c1-110
  done
done
