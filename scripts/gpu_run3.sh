set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|kde selections|Error|error" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench exit $?"; cat gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_r1c.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 3 -c 1 -f -o gpurun_out/prof_forward_r1c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1; echo "ncu2 exit $?"
