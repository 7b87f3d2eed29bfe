mkdir -p gpurun_out/r02_c5
timeout 1500 python scripts/config5_run.py --N 5000 --chunks 20 --snps-per-chunk 50000 --memory 1.74 --max-live 4 --cpu-snps 600 --keep-log gpurun_out/r02_c5/config5_n5000.json > gpurun_out/r02_c5/config5_n5000.log 2>&1; echo "config5 rc=$?"
tail -8 gpurun_out/r02_c5/config5_n5000.log | cut -c1-900
free -g | head -2
