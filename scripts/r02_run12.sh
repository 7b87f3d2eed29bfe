mkdir -p gpurun_out/r02l
timeout 900 python -m pytest tests/test_paint_gpu.py -m gpu -x -q > gpurun_out/r02l/t.log 2>&1; tail -2 gpurun_out/r02l/t.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l/bench_n1.json 2> gpurun_out/r02l/bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02l/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['prep_ms'], d['roofline']['frac_of_nominal_issue'], d['e2e']['value'], d['e2e']['runs_ms'], d['gpu_launches'])
"
