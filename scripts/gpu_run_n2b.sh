set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 scripts/check_sharded.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -30
