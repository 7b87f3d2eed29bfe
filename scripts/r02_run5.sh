mkdir -p gpurun_out/r02e
ncu --set full --import-source on --clock-control none -k regex:repaint_kernel -s 1 -c 1 -f -o gpurun_out/r02e/repaint_ck4 python scripts/prof_window.py 1000 50000 > gpurun_out/r02e/ncu.log 2>&1
tail -3 gpurun_out/r02e/ncu.log
timeout 600 python -m pytest tests/test_paint_gpu.py -m gpu -x -q -k "hapbits or duplicate" > gpurun_out/r02e/t.log 2>&1; tail -3 gpurun_out/r02e/t.log
