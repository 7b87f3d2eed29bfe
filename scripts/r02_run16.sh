mkdir -p gpurun_out/r02p
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02p/gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p/gputests_2gpu.log
tail -3 gpurun_out/r02p/gputests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02p/bench_n2.json 2> gpurun_out/r02p/bench_n2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02p/bench_n1.json 2> gpurun_out/r02p/bench_n1.err; echo "bench1 rc=$?"
python -c "
import json
for f in ('bench_n1','bench_n2'):
    d=json.loads(open('gpurun_out/r02p/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['runs_ms'], d['roofline']['frac_of_nominal_issue'], d['roofline']['traffic'], d['roofline'].get('traffic_source'))
    s=d['sharded']; print('  sharded', s['ms_stage'], s['ms_paint_max'], s['kernel_frac_nominal'], s.get('files_identical_to_1gpu'), s['dij']['ok'])
"
