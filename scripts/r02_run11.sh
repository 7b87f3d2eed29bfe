mkdir -p gpurun_out/r02k
timeout 900 python -m pytest tests/test_paint_gpu.py -m gpu -x -q > gpurun_out/r02k/t.log 2>&1; tail -2 gpurun_out/r02k/t.log
RELATE_BENCH_CONFIG4=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02k/bench_c4.json 2> gpurun_out/r02k/bench_c4.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02k/bench_c4.json').read().strip().splitlines()[-1])
print(d['sharded_config4'].get('e2e_resident_1gpu'))
print(d['value'], d['e2e']['value'], d['e2e']['runs_ms'], d['e2e_resident']['runs_ms'], d['window_repaint']['repaint_kernel_ms'])
"
