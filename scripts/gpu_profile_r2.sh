# usage: bash scripts/gpu_profile_r2.sh TAG -- the evidence of one round: launch lists, ncu --set full captures of the heavy kernels, bench lines
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:forward_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_forward_tc_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwdtc.log 2>&1; echo "ncu fwd exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kde_screened -s 3 -c 1 -f -o gpurun_out/prof_kde_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kde.log 2>&1; echo "ncu kde exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:median_overlap -s 2 -c 1 -f -o gpurun_out/prof_median_$TAG python scripts/eucl_once.py 4 > gpurun_out/ncu_med.log 2>&1; echo "ncu median exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dtw_fixed -s 2 -c 1 -f -o gpurun_out/prof_dtw_$TAG python scripts/eucl_once.py 4 > gpurun_out/ncu_dtw.log 2>&1; echo "ncu dtw exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_eucl_$TAG.csv python scripts/eucl_once.py 4 > gpurun_out/ncu_eucl.log 2>&1; echo "ncu eucl exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
timeout 300 python scripts/config_timings.py > gpurun_out/config_timings_$TAG.json 2> gpurun_out/config_timings_$TAG.err; cut -c1-400 gpurun_out/config_timings_$TAG.json
