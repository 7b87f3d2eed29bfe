mkdir -p gpurun_out/r02h
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02h/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h/gputests.log
tail -4 gpurun_out/r02h/gputests.log
grep -n "gate\|d_ij lens\|64 targets" gpurun_out/r02h/gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02h/bench_n1.json 2> gpurun_out/r02h/bench_n1.err; echo "bench1 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02h/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e_resident'])
"
