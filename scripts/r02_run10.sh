mkdir -p gpurun_out/r02j
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02j/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j/gputests.log
tail -3 gpurun_out/r02j/gputests.log
grep -n "gate:\|d_ij lens:\|64 targets:" gpurun_out/r02j/gputests.log
RELATE_BENCH_CONFIG4=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j/bench_c4.json 2> gpurun_out/r02j/bench_c4.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02j/bench_c4.json').read().strip().splitlines()[-1])
for k in ('sharded','sharded_config4'):
    s=d[k]; print(k, s['ms_stage'], s['ms_stage_runs'], s['ms_paint_max'], s['kernel_frac_nominal'], s['breakdown_ms'])
print(d['sharded_config4'].get('e2e_resident_1gpu'))
print(d['value'], d['e2e']['value'], d['e2e']['runs_ms'], d['e2e_resident']['runs_ms'], d['window_repaint']['repaint_kernel_ms'])
"
