set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s -x -k "tensor_core_forward or fused_forward or module_forwards or materialised" > gpurun_out/tc3.log 2>&1; echo "tc3 exit $?" >> gpurun_out/tc3.log
grep -E "passed|failed|FAILED|Error|error" gpurun_out/tc3.log | tail
timeout 300 python scripts/tc_cycles.py
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); print(d['value'], d['ms_per_step'], d['kernels_ms'])"; tail -3 gpurun_out/bench_r1f.err
