set -x
timeout 120 python scripts/tc_cycles.py 2>&1 | tail -24
