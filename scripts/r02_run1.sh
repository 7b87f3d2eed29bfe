mkdir -p gpurun_out/r02a
(df -T /tmp /dev/shm; nproc; free -g; nvidia-smi -L) > gpurun_out/r02a/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a/gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/gputests_2gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a/bench_n1.json 2> gpurun_out/r02a/bench_n1.err; echo "bench1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02a/bench_n2.json 2> gpurun_out/r02a/bench_n2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a/bench_ref.json 2> gpurun_out/r02a/bench_ref.err; echo "ref rc=$?"
tail -5 gpurun_out/r02a/gputests_2gpu.log
