mkdir -p gpurun_out/r02c
cd /root/repo
for nk in 148 296 592 1000; do
  for plain in 0 1; do
    python scripts/prof_case.py 1000 50000 4 0 0 $nk 0 1 0 $plain | sed "s/^/plain=$plain /"
  done
done > gpurun_out/r02c/lone.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:paint_look -s 2 -c 1 -f -o gpurun_out/r02c/look python scripts/prof_case.py 1000 50000 4 > gpurun_out/r02c/ncu_look.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:paint_kernel -s 2 -c 1 -f -o gpurun_out/r02c/plain python scripts/prof_case.py 1000 50000 4 0 0 1000 0 0 0 1 > gpurun_out/r02c/ncu_plain.log 2>&1
cat gpurun_out/r02c/lone.txt
ls -la gpurun_out/r02c
