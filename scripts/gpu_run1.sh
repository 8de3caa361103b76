set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; echo "bench exit $?"; cat gpurun_out/bench_r1a.json; tail -5 gpurun_out/bench_r1a.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider -k "edge_n65 or edge_n300 or median or mobius_linear or reference_utils" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?"; tail -15 gpurun_out/sanitizer.log
