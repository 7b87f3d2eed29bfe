mkdir -p gpurun_out/r02m
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02m/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:paint_kernel -s 2 -c 1 -f -o gpurun_out/r02m/paint_v6 python scripts/prof_case.py 1000 50000 4 > gpurun_out/r02m/ncu_paint.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:paint_kernel -s 1 -c 1 -f -o gpurun_out/r02m/paint_n10k python scripts/prof_case.py 10000 20000 3 0 0 2236 > gpurun_out/r02m/ncu_n10k.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:repaint_kernel -s 1 -c 1 -f -o gpurun_out/r02m/repaint_ck4 python scripts/prof_window.py 1000 50000 > gpurun_out/r02m/ncu_repaint.log 2>&1
ls -la gpurun_out/r02m
