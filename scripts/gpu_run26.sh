set -x
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
