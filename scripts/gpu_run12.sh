set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "kde" 2>&1 | tail -4
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1h.json')); print(d['value'], d['ms_per_step'], d['kernels_ms'])"; tail -3 gpurun_out/bench_r1h.err
