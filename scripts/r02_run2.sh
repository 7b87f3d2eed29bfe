mkdir -p gpurun_out/r02b
timeout 600 python scripts/ab_look.py > gpurun_out/r02b/ab_look.txt 2>&1; echo "ab rc=$?"
timeout 1200 python -m pytest tests/test_paint_gpu.py -m gpu -x -q -s > gpurun_out/r02b/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b/gputests.log
tail -15 gpurun_out/r02b/gputests.log
cat gpurun_out/r02b/ab_look.txt
