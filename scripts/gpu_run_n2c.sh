set -x
timeout 240 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 scripts/check_sharded.py 2>&1 | grep -E "sharded ==|SHARDED_CHECK|Error"
