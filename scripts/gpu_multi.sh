# usage: bash scripts/gpu_multi.sh N TAG   -- sharded == single-GPU check, bench.py and the cfg4 script on N GPUs of this box
set -x
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/check_sharded.py 2>&1 | grep -v "^\[W\|^W\|^\*\*\*" | tail -12
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench exit $?"; python - <<P
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","e2e","kernels_ms","gpu_launches")})
P
tail -3 gpurun_out/bench_${TAG}_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/cfg4_multivariate_sharded.py --steps 5 > gpurun_out/cfg4_${TAG}_n$N.json 2> gpurun_out/cfg4_${TAG}_n$N.err; echo "cfg4 exit $?"; cat gpurun_out/cfg4_${TAG}_n$N.json; tail -3 gpurun_out/cfg4_${TAG}_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 scripts/sharded_phases.py 2> gpurun_out/phases_${TAG}_n$N.err | tee gpurun_out/phases_${TAG}_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 scripts/cfg5_sweep_sharded.py 2> gpurun_out/cfg5_${TAG}_n$N.err | tee gpurun_out/cfg5_${TAG}_n$N.json; tail -2 gpurun_out/cfg5_${TAG}_n$N.err
HYPAD_PEER_EXCHANGE=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n${N}_nccl.json 2> gpurun_out/bench_${TAG}_n${N}_nccl.err; echo "bench (NCCL exchanges) exit $?"; python - <<P
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n${N}_nccl.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","stage_exchange")}, d["e2e"]["ms_per_step"])
P
HYPAD_PEER_EXCHANGE=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 scripts/sharded_phases.py 2> /dev/null | tee gpurun_out/phases_${TAG}_n${N}_nccl.json
