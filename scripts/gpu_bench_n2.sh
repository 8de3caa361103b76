set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 200 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -12
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1r_n2.json 2> gpurun_out/bench_r1r_n2.err; echo "bench2 exit $?"; cat gpurun_out/bench_r1r_n2.json | cut -c1-400; tail -3 gpurun_out/bench_r1r_n2.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/cfg4_multivariate_sharded.py --steps 5 > gpurun_out/cfg4_r1r_n2.json 2> gpurun_out/cfg4_r1r_n2.err; echo "cfg4 exit $?"; cat gpurun_out/cfg4_r1r_n2.json; tail -3 gpurun_out/cfg4_r1r_n2.err
