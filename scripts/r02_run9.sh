mkdir -p gpurun_out/r02i
timeout 900 python -m pytest tests/test_parity_at_size_gpu.py tests/test_paint_gpu.py -m gpu -x -q -k "config2 or resident or window or hapbits or mode_all" > gpurun_out/r02i/t.log 2>&1; tail -3 gpurun_out/r02i/t.log
timeout 600 python scripts/config5_run.py --N 500 --chunks 4 --snps-per-chunk 25000 --cpu-snps 2000 --keep-log gpurun_out/r02i/config5_small.json > gpurun_out/r02i/config5_small.log 2>&1; echo "config5 rc=$?"; tail -3 gpurun_out/r02i/config5_small.log | cut -c1-1500
RELATE_BENCH_CONFIG4=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02i/bench_c4.json 2> gpurun_out/r02i/bench_c4.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02i/bench_c4.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('sharded_config4'),indent=1)[:6000])
print(d['e2e_resident'])
"
