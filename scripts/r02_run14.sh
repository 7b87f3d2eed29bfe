mkdir -p gpurun_out/r02n
( for shape in "1000 50000" "600 30000" "2000 20000"; do echo "shape $shape"; python scripts/prof_window.py $shape; done ) > gpurun_out/r02n/window.txt 2>&1
cat gpurun_out/r02n/window.txt
timeout 600 python -m pytest tests/test_paint_gpu.py -m gpu -x -q -k "window or resident" 2>&1 | tail -2
