set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|kde selections|Error|error" gpurun_out/pytest_gpu.log | tail -24
timeout 300 python scripts/tc_cycles.py
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1g.json')); print(d['value'], d['ms_per_step'], d['kernels_ms'])"; tail -3 gpurun_out/bench_r1g.err
