set -x
nvidia-smi --query-gpu=index,name --format=csv | wc -l
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r1n_n8.json 2> gpurun_out/bench_r1n_n8.err; echo "bench8 exit $?"; wc -l gpurun_out/bench_r1n_n8.json; cut -c1-300 gpurun_out/bench_r1n_n8.json; tail -3 gpurun_out/bench_r1n_n8.err
