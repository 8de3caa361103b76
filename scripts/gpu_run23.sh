set -x
timeout 300 python scripts/host_tail.py 2>&1 | tail -5
