set -x
timeout 300 python scripts/tc_cycles.py 2>&1 | tail -24
