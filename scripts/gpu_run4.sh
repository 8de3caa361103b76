set -x
timeout 300 python -m pytest tests/test_gpu_tensor_probe.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log
tail -40 gpurun_out/probe.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s -k "end_to_end" > gpurun_out/pytest_e2e.log 2>&1; grep -E "passed|failed|FAILED|conditioning|kde selections" gpurun_out/pytest_e2e.log | tail -30
