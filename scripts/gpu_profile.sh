set -x
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:forward_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_forward_tc_r1n python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwdtc.log 2>&1; echo "ncu2 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kde_screened -s 3 -c 1 -f -o gpurun_out/prof_kde_r1n python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kde.log 2>&1; echo "ncu3 exit $?"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; echo "bench exit $?"; cat gpurun_out/bench_r1n.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1n.json 2> gpurun_out/bench_ref_r1n.err; echo "ref exit $?"; cat gpurun_out/bench_ref_r1n.json | cut -c1-300
