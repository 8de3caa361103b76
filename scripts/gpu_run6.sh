set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s -x -k "tensor_core_forward" > gpurun_out/tc1.log 2>&1; echo "tc1 exit $?" >> gpurun_out/tc1.log
tail -30 gpurun_out/tc1.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|kde selections|Error|error|max \|tensor" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench exit $?"; cat gpurun_out/bench_r1d.json; tail -3 gpurun_out/bench_r1d.err
