set -x
nvidia-smi --query-gpu=index,name --format=csv | head -6
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r1l_n4.json 2> gpurun_out/bench_r1l_n4.err; echo "bench4 exit $?"; cat gpurun_out/bench_r1l_n4.json | cut -c1-400; tail -3 gpurun_out/bench_r1l_n4.err
