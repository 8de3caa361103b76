timeout 300 python -m pytest tests/test_gpu_tensor_probe.py -m gpu -q --tb=short -p no:cacheprovider -s -k accuracy > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log
tail -40 gpurun_out/probe.log
