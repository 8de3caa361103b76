set -x
timeout 240 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -2
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1m_n2.json 2> gpurun_out/bench_r1m_n2.err; echo "bench2 exit $?"; wc -l gpurun_out/bench_r1m_n2.json; cut -c1-200 gpurun_out/bench_r1m_n2.json
