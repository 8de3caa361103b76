set -x
timeout 240 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -6
timeout 120 python scripts/tc_cycles.py 2>&1 | tail -24
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_cur.json')); print(d['value'], d['ms_per_step'], d['kernels_ms'])"; tail -3 gpurun_out/bench_cur.err
