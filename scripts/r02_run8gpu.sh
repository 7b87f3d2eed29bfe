mkdir -p gpurun_out/r02_8gpu
nvidia-smi -L > gpurun_out/r02_8gpu/box.txt; nproc >> gpurun_out/r02_8gpu/box.txt; free -g >> gpurun_out/r02_8gpu/box.txt; df -h /tmp /dev/shm >> gpurun_out/r02_8gpu/box.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_8gpu/bench_n8.json 2> gpurun_out/r02_8gpu/bench_n8.err; echo "bench8 rc=$?"
timeout 600 python -m pytest tests/test_paint_gpu.py -m gpu -x -q -s -k "multi_gpu or small_batches or config5_shape or hapbits" > gpurun_out/r02_8gpu/multigpu_tests.log 2>&1; tail -2 gpurun_out/r02_8gpu/multigpu_tests.log
timeout 900 python scripts/config5_run.py --N 2000 --chunks 20 --snps-per-chunk 50000 --keep-log gpurun_out/r02_8gpu/config5_n2000.json > gpurun_out/r02_8gpu/config5_n2000.log 2>&1; echo "config5 rc=$?"
tail -4 gpurun_out/r02_8gpu/config5_n2000.log | cut -c1-600
python -c "
import json
d=json.loads(open('gpurun_out/r02_8gpu/bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['runs_ms'], d['e2e']['breakdown_ms'])
for k in ('sharded','sharded_config4'):
    s=d.get(k,{}); print(k, {x:s.get(x) for x in ('ms_stage','ms_stage_runs','ms_paint_max','kernel_frac_nominal','files_identical_to_1gpu','ms_stage_1gpu_same_box','breakdown_ms','dij','error','skipped','e2e_resident_1gpu')})
"
