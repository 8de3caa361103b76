set -x
timeout 300 python scripts/finish_scaling.py 2>&1 | tail -6
