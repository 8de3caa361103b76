set -x
timeout 240 python scripts/config_timings.py > gpurun_out/config_timings.json 2> gpurun_out/config_timings.err; echo "configs exit $?"; cat gpurun_out/config_timings.json; tail -3 gpurun_out/config_timings.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:pw_distance_kernel -s 4 -c 1 -f -o gpurun_out/prof_pw_distance_r1o python scripts/next_rows_timing.py > gpurun_out/ncu_pw.log 2>&1; echo "ncu pw exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:segments_aggregate_kernel -s 4 -c 1 -f -o gpurun_out/prof_segments_aggregate_r1o python scripts/next_rows_timing.py > gpurun_out/ncu_agg.log 2>&1; echo "ncu agg exit $?"
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_r1o.json; tail -2 gpurun_out/bench_r1o.err
timeout 200 python scripts/cfg4_multivariate_sharded.py --steps 5 > gpurun_out/cfg4_n1.json 2> gpurun_out/cfg4_n1.err; echo "cfg4 exit $?"; cat gpurun_out/cfg4_n1.json; tail -2 gpurun_out/cfg4_n1.err
timeout 90 python scripts/next_rows_timing.py > gpurun_out/next_rows_timing_r1o.json 2> gpurun_out/next_rows_timing_r1o.err; echo "timing exit $?"; cat gpurun_out/next_rows_timing_r1o.json
