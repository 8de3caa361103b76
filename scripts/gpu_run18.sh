set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kde_screened -s 3 -c 1 -f -o gpurun_out/prof_kde_cur python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kde.log 2>&1; echo "ncu exit $?"
