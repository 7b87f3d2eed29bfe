mkdir -p gpurun_out/r02d
timeout 1200 python -m pytest tests/test_paint_gpu.py -m gpu -x -q -s > gpurun_out/r02d/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d/gputests.log
tail -6 gpurun_out/r02d/gputests.log
( for ck in 4 8; do for shape in "1000 50000" "600 30000"; do echo "CK=$ck $shape"; RP_REPAINT_CK=$ck python scripts/prof_window.py $shape; done; done
  for shape in "2000 20000" "5000 20000"; do echo "shape $shape"; python scripts/prof_window.py $shape; done ) > gpurun_out/r02d/window.txt 2>&1
cat gpurun_out/r02d/window.txt
( bash scripts/sweep_variants.sh spin256 spin1024; for nk in 296 888 1000; do python scripts/prof_case.py 1000 50000 3 0 0 $nk | tail -1 | sed "s/^/base /" | cut -c1-110; done ) > gpurun_out/r02d/spin.txt 2>&1
cat gpurun_out/r02d/spin.txt
