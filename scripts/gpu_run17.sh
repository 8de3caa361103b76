set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_forward_tc_cur python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwdtc.log 2>&1; echo "ncu exit $?"
