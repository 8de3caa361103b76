set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 240 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -12
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1l_n2.json 2> gpurun_out/bench_r1l_n2.err; echo "bench2 exit $?"; cat gpurun_out/bench_r1l_n2.json | cut -c1-600; tail -5 gpurun_out/bench_r1l_n2.err
timeout 240 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1l_n1b.json 2> gpurun_out/bench_r1l_n1b.err; echo "bench1 exit $?"; cat gpurun_out/bench_r1l_n1b.json | cut -c1-400
