mkdir -p gpurun_out/r02f
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02f/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f/gputests.log
tail -5 gpurun_out/r02f/gputests.log
grep -n "gate\|d_ij lens\|64 targets" gpurun_out/r02f/gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f/bench_n1.json 2> gpurun_out/r02f/bench_n1.err; echo "bench1 rc=$?"
( for shape in "1000 50000" "600 30000" "2000 20000" "5000 20000"; do echo "shape $shape"; python scripts/prof_window.py $shape; done ) > gpurun_out/r02f/window.txt 2>&1
cat gpurun_out/r02f/window.txt
