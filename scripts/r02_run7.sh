mkdir -p gpurun_out/r02g
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02g/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g/gputests.log
tail -4 gpurun_out/r02g/gputests.log
grep -n "gate\|d_ij lens\|64 targets" gpurun_out/r02g/gputests.log
RELATE_PAINT_LIB=$PWD/variants/lib_tma.so timeout 600 python -m pytest tests/test_paint_gpu.py -m gpu -x -q -k "test_fp32_matches_oracle or parked or config4_shape" > gpurun_out/r02g/tma_tests.log 2>&1; tail -2 gpurun_out/r02g/tma_tests.log
( for v in base tma; do
    for shape in "5000 20000 2 0 0 5000" "10000 20000 2 0 0 2236" "3000 20000 2 0 0 3000"; do
      if [ $v = base ]; then python scripts/prof_case.py $shape | tail -1 | sed "s/^/base /" | cut -c1-170;
      else RELATE_PAINT_LIB=$PWD/variants/lib_tma.so python scripts/prof_case.py $shape | tail -1 | sed "s/^/tma  /" | cut -c1-170; fi
    done
  done ) > gpurun_out/r02g/tma.txt 2>&1
cat gpurun_out/r02g/tma.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g/bench_n1.json 2> gpurun_out/r02g/bench_n1.err; echo "bench1 rc=$?"
